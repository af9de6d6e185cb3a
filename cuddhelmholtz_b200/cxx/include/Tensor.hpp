// Column-major N-d views and owning tensors with the reference's interface (include/Tensor.hpp): first index fastest,
// usable on host and device, implicit conversion to the raw pointer, reshape() helpers and the dvec/dmat/ivec/...
// typedefs. Views are trivially copyable (no vtable), so they can be captured by value in device lambdas.
#ifndef TENSOR_HPP
#define TENSOR_HPP

#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "cuddh_config.hpp"
#include "cuddh_error.hpp"

namespace cuddh
{
    namespace detail
    {
        template <int Dim>
        struct Extents
        {
            int n[Dim];

            template <typename... Sizes>
            __host__ __device__ int set(Sizes... sizes)
            {
                static_assert(sizeof...(sizes) == Dim, "Wrong number of dimensions specified.");
                const int s[] = {static_cast<int>(sizes)...};
                int total = 1;
                for (int d = 0; d < Dim; ++d) {
                    if (s[d] < 0)
                        cuddh_error("Tensor error: tensor cannot have negative dimensions.");
                    n[d] = s[d];
                    total *= s[d];
                }
                return total;
            }

            template <typename... Ids>
            __host__ __device__ int offset(Ids... ids) const
            {
                static_assert(sizeof...(ids) == Dim, "Wrong number of indices specified.");
                const int k[] = {static_cast<int>(ids)...};
                int off = 0;
                for (int d = Dim - 1; d >= 0; --d) {
#ifdef CUDDH_DEBUG
                    if (k[d] < 0 || k[d] >= n[d])
                        cuddh_error("Tensor error: tensor index out of range.");
#endif
                    off = k[d] + n[d] * off;
                }
                return off;
            }
        };
    } // namespace detail

    /// read/write view of an externally managed array with N-d column-major indexing
    template <int Dim, typename scalar>
    class TensorWrapper
    {
    protected:
        detail::Extents<Dim> ext;
        int len;
        scalar * ptr;

    public:
        __host__ __device__ TensorWrapper() : ext{}, len(0), ptr(nullptr) {}

        template <typename... Sizes>
        __host__ __device__ explicit TensorWrapper(scalar * data_, Sizes... shape_) : ptr(data_)
        {
            static_assert(Dim > 0, "Tensor must have a positive number of dimensions");
            len = ext.set(shape_...);
        }

        template <typename... Ids> __host__ __device__ scalar & at(Ids... ids) { return ptr[ext.offset(ids...)]; }
        template <typename... Ids> __host__ __device__ const scalar & at(Ids... ids) const { return ptr[ext.offset(ids...)]; }
        template <typename... Ids> __host__ __device__ scalar & operator()(Ids... ids) { return ptr[ext.offset(ids...)]; }
        template <typename... Ids> __host__ __device__ const scalar & operator()(Ids... ids) const { return ptr[ext.offset(ids...)]; }

        __host__ __device__ scalar & operator[](int idx) { return ptr[idx]; }
        __host__ __device__ const scalar & operator[](int idx) const { return ptr[idx]; }

        __host__ __device__ operator scalar *() { return ptr; }
        __host__ __device__ operator const scalar *() const { return ptr; }

        __host__ __device__ scalar * data() { return ptr; }
        __host__ __device__ const scalar * data() const { return ptr; }
        __host__ __device__ scalar * begin() { return ptr; }
        __host__ __device__ scalar * end() { return ptr + len; }
        __host__ __device__ const scalar * begin() const { return ptr; }
        __host__ __device__ const scalar * end() const { return ptr + len; }

        __host__ __device__ const int * shape() const { return ext.n; }
        __host__ __device__ int shape(int d) const { return ext.n[d]; }
        __host__ __device__ int size() const { return len; }
    };

    template <typename scalar, typename... Sizes>
    __host__ __device__ inline TensorWrapper<sizeof...(Sizes), scalar> reshape(scalar * data, Sizes... shape)
    {
        return TensorWrapper<sizeof...(Sizes), scalar>(data, shape...);
    }

    template <typename scalar, int Dim, typename... Sizes>
    __host__ __device__ inline TensorWrapper<sizeof...(Sizes), scalar> reshape(TensorWrapper<Dim, scalar> tensor, Sizes... shape)
    {
        return TensorWrapper<sizeof...(Sizes), scalar>(tensor.data(), shape...);
    }

    /// owning tensor (host memory, zero-initialised)
    template <int Dim, typename scalar>
    class Tensor : public TensorWrapper<Dim, scalar>
    {
        std::vector<scalar> mem;

        void rebind() { this->ptr = mem.empty() ? nullptr : mem.data(); }

    public:
        Tensor() : TensorWrapper<Dim, scalar>() {}

        template <typename... Sizes>
        explicit Tensor(Sizes... shape_) : TensorWrapper<Dim, scalar>(nullptr, shape_...), mem((size_t)this->len, scalar())
        {
            rebind();
        }

        Tensor(const Tensor & t) : TensorWrapper<Dim, scalar>(t), mem(t.mem) { rebind(); }
        Tensor & operator=(const Tensor & t)
        {
            if (this != &t) {
                TensorWrapper<Dim, scalar>::operator=(t);
                mem = t.mem;
                rebind();
            }
            return *this;
        }
        Tensor(Tensor && t) noexcept : TensorWrapper<Dim, scalar>(t), mem(std::move(t.mem))
        {
            rebind();
            t.ptr = nullptr;
            t.len = 0;
        }
        Tensor & operator=(Tensor && t) noexcept
        {
            TensorWrapper<Dim, scalar>::operator=(t);
            mem = std::move(t.mem);
            rebind();
            t.ptr = nullptr;
            t.len = 0;
            return *this;
        }

        /// change the shape; memory is reallocated (and zeroed) only when it has to grow
        template <typename... Sizes>
        void reshape(Sizes... shape_)
        {
            const int new_len = this->ext.set(shape_...);
            if ((size_t)new_len > mem.size()) {
                mem.assign((size_t)new_len, scalar());
                rebind();
            }
            this->len = new_len;
        }
    };

    template <typename scalar> using VectorWrapper = TensorWrapper<1, scalar>;
    template <typename scalar> using MatrixWrapper = TensorWrapper<2, scalar>;
    template <typename scalar> using CubeWrapper = TensorWrapper<3, scalar>;

    typedef TensorWrapper<1, double> dvec_wrapper;
    typedef TensorWrapper<1, const double> const_dvec_wrapper;
    typedef TensorWrapper<2, double> dmat_wrapper;
    typedef TensorWrapper<2, const double> const_dmat_wrapper;
    typedef TensorWrapper<3, double> dcube_wrapper;
    typedef TensorWrapper<3, const double> const_dcube_wrapper;
    typedef TensorWrapper<1, int> ivec_wrapper;
    typedef TensorWrapper<1, const int> const_ivec_wrapper;
    typedef TensorWrapper<2, int> imat_wrapper;
    typedef TensorWrapper<2, const int> const_imat_wrapper;
    typedef TensorWrapper<3, int> icube_wrapper;
    typedef TensorWrapper<3, const int> const_icube_wrapper;

    template <typename scalar> using Vec = Tensor<1, scalar>;
    template <typename scalar> using Matrix = Tensor<2, scalar>;
    template <typename scalar> using Cube = Tensor<3, scalar>;

    typedef Vec<double> dvec;
    typedef Matrix<double> dmat;
    typedef Cube<double> dcube;
    typedef Vec<int> ivec;
    typedef Matrix<int> imat;
    typedef Cube<int> icube;
} // namespace cuddh

#endif
