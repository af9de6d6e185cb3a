// forall / forall_{1,2,3}d launch helpers with the reference's call signatures (include/forall.hpp:22-64): user code
// and the header templates of this layer pass extended __device__ lambdas. forall uses a grid-stride loop so that any
// n fits one launch; the block forms run one CTA per work item with the requested block shape.
#ifndef CUDDH_FORALL_HPP
#define CUDDH_FORALL_HPP

#include <cuda_runtime.h>

#ifndef CUDDH_FORALL_BLOCK_SIZE
#define CUDDH_FORALL_BLOCK_SIZE 256
#endif

namespace cuddh
{
    namespace detail
    {
        template <typename F>
        __global__ void each_index(int n, F fun)
        {
            for (long long k = threadIdx.x + (long long)blockIdx.x * blockDim.x; k < n; k += (long long)gridDim.x * blockDim.x)
                fun((int)k);
        }

        template <typename F>
        __global__ void each_block(int n, F fun)
        {
            if ((int)blockIdx.x < n)
                fun((int)blockIdx.x);
        }

        inline unsigned grid_for(long long items) { return (unsigned)(items < 1 ? 1 : (items > 1048576 ? 1048576 : items)); }
    } // namespace detail

    template <typename LAMBDA>
    inline void forall(int n, LAMBDA && fun)
    {
        if (n <= 0) return;
        const long long blocks = ((long long)n + CUDDH_FORALL_BLOCK_SIZE - 1) / CUDDH_FORALL_BLOCK_SIZE;
        detail::each_index<<<detail::grid_for(blocks), CUDDH_FORALL_BLOCK_SIZE>>>(n, fun);
    }

    template <typename LAMBDA>
    inline void forall_1d(int bx, int n, LAMBDA && fun)
    {
        if (n <= 0) return;
        detail::each_block<<<(unsigned)n, bx>>>(n, fun);
    }

    template <typename LAMBDA>
    inline void forall_2d(int bx, int by, int n, LAMBDA && fun)
    {
        if (n <= 0) return;
        detail::each_block<<<(unsigned)n, dim3(bx, by)>>>(n, fun);
    }

    template <typename LAMBDA>
    inline void forall_3d(int bx, int by, int bz, int n, LAMBDA && fun)
    {
        if (n <= 0) return;
        detail::each_block<<<(unsigned)n, dim3(bx, by, bz)>>>(n, fun);
    }
} // namespace cuddh

#endif
