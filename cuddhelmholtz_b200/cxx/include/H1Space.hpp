// H1Space and FaceSpace (reference include/H1Space.hpp): global DOF numbering of the tensor-product H1 space and the
// subspace of DOFs on a set of faces. Index data is built by the library with the reference's numbering.
#ifndef CUDDH_H1_SPACE_HPP
#define CUDDH_H1_SPACE_HPP

#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>

#include "Basis.hpp"
#include "HostDeviceArray.hpp"
#include "Mesh2D.hpp"
#include "Operator.hpp"
#include "Tensor.hpp"
#include "forall.hpp"

namespace cuddh
{
    class H1Space
    {
    public:
        H1Space(const Mesh2D & mesh, const Basis & basis) : n_elem(mesh.n_elem()), n_basis(basis.size()), _mesh(mesh), _basis(basis)
        {
            cuddh_h1space_t raw = nullptr;
            cuddh_check(cuddh_b200_h1space_create(mesh.handle(), n_basis, &raw));
            h.reset(raw, [](cuddh_h1space_t p) { cuddh_b200_h1space_destroy(p); });
            ndof = (int)cuddh_b200_h1space_size(raw);
        }

        int size() const { return ndof; }

        /// (n_basis, n_basis, n_elem): global index of the (i, j) node of element el
        const_icube_wrapper global_indices(MemorySpace m) const
        {
            const int * p = (m == MemorySpace::HOST) ? cuddh_b200_h1space_host_indices(h.get()) : cuddh_b200_h1space_device_indices(h.get());
            return reshape(p, n_basis, n_basis, n_elem);
        }

        const Mesh2D & mesh() const { return _mesh; }
        const Basis & basis() const { return _basis; }

        /// (2, ndof) collocation points of the nodal DOFs
        const_dmat_wrapper physical_coordinates(MemorySpace m) const
        {
            const double * p = (m == MemorySpace::HOST) ? cuddh_b200_h1space_host_coordinates(h.get()) : cuddh_b200_h1space_device_coordinates(h.get());
            return reshape(p, 2, ndof);
        }

        cuddh_h1space_t handle() const { return h.get(); }

    private:
        const int n_elem;
        const int n_basis;
        const Mesh2D & _mesh;
        const Basis & _basis;
        int ndof;
        std::shared_ptr<cuddh_h1space_s> h;
    };

    class FaceSpace
    {
    public:
        FaceSpace(const H1Space & fem_, int n_faces, const int * faces_) : fem(fem_), _n_faces(n_faces), n_basis(fem_.basis().size())
        {
            cuddh_facespace_t raw = nullptr;
            cuddh_check(cuddh_b200_facespace_create(fem.handle(), n_faces, faces_, &raw));
            h.reset(raw, [](cuddh_facespace_t p) { cuddh_b200_facespace_destroy(p); });
            ndof = (int)cuddh_b200_facespace_size(raw);
        }

        int size() const { return ndof; }
        int n_faces() const { return _n_faces; }

        const_ivec_wrapper faces(MemorySpace m) const { return reshape(ptr(2, m), _n_faces); }
        /// (n_basis, n_faces): face-space index of the i-th node of face f
        const_imat_wrapper subspace_indices(MemorySpace m) const { return reshape(ptr(0, m), n_basis, _n_faces); }
        /// (size()): H1 index of each face-space DOF
        const_ivec_wrapper global_indices(MemorySpace m) const { return reshape(ptr(1, m), ndof); }

        void restrict(const double * x, double * y) const { cuddh_check(cuddh_b200_facespace_restrict(h.get(), x, y, nullptr)); } ///< y[i] = x[proj[i]]
        void prolong(const double * x, double * y) const { cuddh_check(cuddh_b200_facespace_prolong(h.get(), x, y, nullptr)); }   ///< y[proj[i]] += x[i]
        void orth(double * x) const { cuddh_check(cuddh_b200_facespace_orth(h.get(), x, nullptr)); }                              ///< x[proj[i]] = 0

        const H1Space & h1_space() const { return fem; }

        const Mesh2D::EdgeMetricCollection & metrics(const QuadratureRule & quad) const
        {
            const std::string key = quad.name();
            auto it = _metrics.find(key);
            if (it == _metrics.end())
                it = _metrics.emplace(key, Mesh2D::EdgeMetricCollection(fem.mesh(), _n_faces, ptr(2, MemorySpace::HOST), quad)).first;
            return it->second;
        }

        cuddh_facespace_t handle() const { return h.get(); }

    private:
        const int * ptr(int which, MemorySpace m) const { return cuddh_b200_facespace_indices_ptr(h.get(), which, m == MemorySpace::DEVICE); }

        const H1Space & fem;
        const int _n_faces;
        const int n_basis;
        int ndof;
        std::shared_ptr<cuddh_facespace_s> h;
        mutable std::unordered_map<std::string, Mesh2D::EdgeMetricCollection> _metrics;
    };
} // namespace cuddh

#endif
