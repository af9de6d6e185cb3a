// BLAS-1 kernels and restarted GMRES (reference include/linalg.hpp, source/linalg.cpp, source/gmres.cpp).
//
// Reference behaviour being replaced: 32-thread reduction CTAs with one same-address atomicAdd per 256
// elements, a cudaMalloc + memset + blocking D2H + cudaFree per dot, and k+1 (dot, axpy) launch pairs with
// k+1 host syncs per Arnoldi step. Here:
//   * reductions are two-stage and deterministic: per-CTA partials in fixed slots, the last CTA to finish
//     (ticket counter) sums them in slot order; grid = a multiple of the SM count, 128-bit loads;
//   * the modified Gram-Schmidt sweep keeps its exact semantics (h_j = <w, v_j> AFTER w has been updated
//     with h_{j-1}) but every step is ONE fused kernel: w -= h_{j-1} v_{j-1} and the partial dot with v_j in
//     the same pass, h_{j-1} read from device memory, so the host syncs once per Arnoldi step instead of
//     k+1 times;
//   * x += sum_k eta_k v_k is one pass (same left-to-right order as the reference's axpby chain).
#include "linalg.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <iomanip>

namespace cb200
{
    namespace
    {
        constexpr int RED_THREADS = 256;
        constexpr int MAX_RED_BLOCKS = 1184; // 8 CTAs x 148 SMs

        int sm_count()
        {
            static int n = 0;
            if (!n) {
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
                if (n <= 0)
                    n = 148;
            }
            return n;
        }

        int red_blocks(int64_t n)
        {
            const int64_t want = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
            return (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min(MAX_RED_BLOCKS, 8 * sm_count())));
        }

        template <typename T>
        __device__ __forceinline__ T warp_sum(T v)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_down_sync(0xffffffffu, v, o);
            return v;
        }

        // block partial -> slot; last block sums slots in order and writes result[0] (optionally sqrt)
        template <typename T>
        __device__ __forceinline__ void block_finish(T acc, T * __restrict__ partials, unsigned * __restrict__ ticket,
                                                     T * __restrict__ result, const bool take_sqrt)
        {
            __shared__ T wsum[RED_THREADS / 32];
            __shared__ bool last;
            acc = warp_sum(acc);
            if ((threadIdx.x & 31) == 0)
                wsum[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                T b = 0;
#pragma unroll
                for (int w = 0; w < RED_THREADS / 32; ++w)
                    b += wsum[w];
                partials[blockIdx.x] = b;
                __threadfence();
                const unsigned t = atomicAdd(ticket, 1u);
                last = (t == gridDim.x - 1);
            }
            __syncthreads();
            if (last) {
                __threadfence();
                T s = 0;
                for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS)
                    s += ((volatile T *)partials)[i];
                // fixed-shape tree over the 256 per-thread sums
                s = warp_sum(s);
                __syncthreads();
                if ((threadIdx.x & 31) == 0)
                    wsum[threadIdx.x >> 5] = s;
                __syncthreads();
                if (threadIdx.x == 0) {
                    T tot = 0;
#pragma unroll
                    for (int w = 0; w < RED_THREADS / 32; ++w)
                        tot += wsum[w];
                    result[0] = take_sqrt ? sqrt(tot) : tot;
                    *ticket = 0;
                }
            }
        }

        template <typename T, int MODE> // MODE 0: x.y   1: (x-y)^2
        __global__ void __launch_bounds__(RED_THREADS)
        reduce_kernel(const int64_t n, const T * __restrict__ x, const T * __restrict__ y, T * __restrict__ partials,
                      unsigned * __restrict__ ticket, T * __restrict__ result, const int take_sqrt)
        {
            T acc = 0;
            const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
            for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n; i += stride) {
                if (MODE == 0)
                    acc = fma(x[i], y[i], acc);
                else {
                    const T e = x[i] - y[i];
                    acc = fma(e, e, acc);
                }
            }
            block_finish(acc, partials, ticket, result, take_sqrt != 0);
        }

        // fused MGS step: w -= h[prev] * vprev (if vprev), then dot(w, vnext) -> h[out] (vnext == null: ||w||)
        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS)
        mgs_step_kernel(const int64_t n, T * __restrict__ w, const T * __restrict__ vprev, const T * __restrict__ vnext,
                        const T * __restrict__ hprev, T * __restrict__ partials, unsigned * __restrict__ ticket,
                        T * __restrict__ hout)
        {
            const T hp = vprev ? hprev[0] : T(0);
            T acc = 0;
            const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
            for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n; i += stride) {
                T wi = w[i];
                if (vprev) {
                    wi = fma(-hp, vprev[i], wi);
                    w[i] = wi;
                }
                acc = fma(wi, vnext ? vnext[i] : wi, acc);
            }
            block_finish(acc, partials, ticket, hout, vnext == nullptr);
        }

        template <typename T>
        __global__ void axpby_kernel(const int64_t n, const T a, const T * __restrict__ x, const T b, T * __restrict__ y)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (i < n)
                y[i] = a * x[i] + b * y[i];
        }
        template <typename T>
        __global__ void scal_kernel(const int64_t n, const T a, T * __restrict__ x)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (i < n)
                x[i] *= a;
        }
        // x <- x / s[0] with s on the device (no host round trip)
        template <typename T>
        __global__ void scal_inv_dev_kernel(const int64_t n, const T * __restrict__ s, T * __restrict__ x)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            const T a = T(1) / s[0];
            if (i < n)
                x[i] *= a;
        }
        template <typename T>
        __global__ void fill_kernel(const int64_t n, const T a, T * __restrict__ x)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (i < n)
                x[i] = a;
        }
        template <typename T>
        __global__ void copy_kernel(const int64_t n, const T * __restrict__ x, T * __restrict__ y)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (i < n)
                y[i] = x[i];
        }
        // x <- x + sum_k eta[k] V[:,k], left to right (gmres.cpp:190-191)
        template <typename T>
        __global__ void multi_axpy_kernel(const int64_t n, const int k1, const T * __restrict__ eta, const T * __restrict__ V,
                                          T * __restrict__ x)
        {
            const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (i >= n)
                return;
            T acc = x[i];
            for (int k = 0; k < k1; ++k)
                acc = eta[k] * V[i + (size_t)k * n] + acc;
            x[i] = acc;
        }

        inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

        // per-thread-of-execution reduction workspace (the reference allocates one per dot call)
        struct RedWork
        {
            DevBuf<double> part_d, res_d;
            DevBuf<float> part_f, res_f;
            DevBuf<unsigned> ticket;
            bool ready = false;
            void init()
            {
                if (ready)
                    return;
                part_d.alloc(MAX_RED_BLOCKS);
                part_f.alloc(MAX_RED_BLOCKS);
                res_d.alloc(64);
                res_f.alloc(64);
                ticket.alloc(1);
                ticket.zero();
                CB_CUDA(cudaDeviceSynchronize());
                ready = true;
            }
            template <typename T> T * part();
            template <typename T> T * res();
        };
        template <> double * RedWork::part<double>() { return part_d.p; }
        template <> float * RedWork::part<float>() { return part_f.p; }
        template <> double * RedWork::res<double>() { return res_d.p; }
        template <> float * RedWork::res<float>() { return res_f.p; }

        RedWork & work()
        {
            static RedWork w;
            w.init();
            return w;
        }
    } // namespace

    template <typename T> void axpby(int64_t n, T a, const T * x, T b, T * y, cudaStream_t s)
    {
        if (n <= 0) return;
        axpby_kernel<T><<<nblk(n), 256, 0, s>>>(n, a, x, b, y);
        CB_LAUNCHED();
    }
    template <typename T> void scal(int64_t n, T a, T * x, cudaStream_t s)
    {
        if (n <= 0) return;
        scal_kernel<T><<<nblk(n), 256, 0, s>>>(n, a, x);
        CB_LAUNCHED();
    }
    template <typename T> void fill(int64_t n, T a, T * x, cudaStream_t s)
    {
        if (n <= 0) return;
        fill_kernel<T><<<nblk(n), 256, 0, s>>>(n, a, x);
        CB_LAUNCHED();
    }
    template <typename T> void copy(int64_t n, const T * x, T * y, cudaStream_t s)
    {
        if (n <= 0) return;
        copy_kernel<T><<<nblk(n), 256, 0, s>>>(n, x, y);
        CB_LAUNCHED();
    }
    template <typename T> T dot(int64_t n, const T * x, const T * y, cudaStream_t s)
    {
        if (n <= 0) return T(0);
        RedWork & w = work();
        reduce_kernel<T, 0><<<red_blocks(n), RED_THREADS, 0, s>>>(n, x, y, w.part<T>(), w.ticket.p, w.res<T>(), 0);
        CB_LAUNCHED();
        T h;
        CB_CUDA(cudaMemcpyAsync(&h, w.res<T>(), sizeof(T), cudaMemcpyDeviceToHost, s));
        CB_CUDA(cudaStreamSynchronize(s));
        return h;
    }
    template <typename T> T dist(int64_t n, const T * x, const T * y, cudaStream_t s)
    {
        if (n <= 0) return T(0);
        RedWork & w = work();
        reduce_kernel<T, 1><<<red_blocks(n), RED_THREADS, 0, s>>>(n, x, y, w.part<T>(), w.ticket.p, w.res<T>(), 0);
        CB_LAUNCHED();
        T h;
        CB_CUDA(cudaMemcpyAsync(&h, w.res<T>(), sizeof(T), cudaMemcpyDeviceToHost, s));
        CB_CUDA(cudaStreamSynchronize(s));
        return std::sqrt(h);
    }

#define CB_INST(T)                                                                         \
    template void axpby<T>(int64_t, T, const T *, T, T *, cudaStream_t);                   \
    template void scal<T>(int64_t, T, T *, cudaStream_t);                                  \
    template void fill<T>(int64_t, T, T *, cudaStream_t);                                  \
    template void copy<T>(int64_t, const T *, T *, cudaStream_t);                          \
    template T dot<T>(int64_t, const T *, const T *, cudaStream_t);                        \
    template T dist<T>(int64_t, const T *, const T *, cudaStream_t);
    CB_INST(double)
    CB_INST(float)
    template void fill<int>(int64_t, int, int *, cudaStream_t);
    template void copy<int>(int64_t, const int *, int *, cudaStream_t);
#undef CB_INST

    // ---------------------------------------------------------------------------------------------
    // GMRES(m): reference source/gmres.cpp:91-235 (t_gmres). Same control flow and stopping rules:
    //   it starts at 1 and runs while it < maxit; inner break on |eta_{k+1}| < tol*||b|| or H(k+1,k) == 0;
    //   true residual recomputed every restart; num_iter = it.
    // ---------------------------------------------------------------------------------------------
    namespace
    {
        template <typename T>
        void givens(T * h, T * cs, T * sn, int k) // gmres.cpp:7-23
        {
            for (int i = 0; i < k; ++i) {
                const T h1 = h[i], h2 = h[i + 1];
                h[i] = cs[i] * h1 + sn[i] * h2;
                h[i + 1] = -sn[i] * h1 + cs[i] * h2;
            }
            const T t = std::hypot(h[k], h[k + 1]);
            cs[k] = h[k] / t;
            sn[k] = h[k + 1] / t;
            h[k] = cs[k] * h[k] + sn[k] * h[k + 1];
            h[k + 1] = 0.0;
        }

        // upper-triangular solve R x = b in place (what dtrsv_/strsv_("U","N","N") compute; gmres.cpp:26-44)
        template <typename T>
        void trsv_upper(int n, const T * R, int ldr, T * b)
        {
            for (int j = n - 1; j >= 0; --j) {
                if (b[j] != T(0)) {
                    b[j] /= R[j + (size_t)ldr * j];
                    const T t = b[j];
                    for (int i = j - 1; i >= 0; --i)
                        b[i] -= t * R[i + (size_t)ldr * j];
                }
            }
        }

        struct Progress // gmres.cpp:46-66
        {
            int it = 0, nt;
            std::string bar;
            explicit Progress(int nt_) : nt(nt_), bar(30, ' ') {}
            void step()
            {
                it = std::min(it + 1, nt - 1);
                bar.at(30 * (it - 1) / nt) = '#';
            }
        };
    } // namespace

    template <typename T>
    GmresResult gmres(int64_t n, T * x, ApplyFn<T> A, void * ctx, const T * b, int m, int maxit, T tol, int verbose,
                      double max_seconds, cudaStream_t s)
    {
        GmresResult out;
        out.success = false;
        out.num_iter = 0;
        out.num_matvec = 0;
        CB_REQUIRE(m >= 1, "gmres: m must be >= 1");
        const T one = 1, zero = 0;
        RedWork & wk = work();
        const int m1 = m + 1;

        const T bnrm = std::sqrt(dot<T>(n, b, b, s));

        DevBuf<T> r_((size_t)n), V_((size_t)n * m1), hdev((size_t)m1 + 2), etadev((size_t)m1);
        T * r = r_.p;
        T * V = V_.p;
        CB_CUDA(cudaMemsetAsync(V, 0, sizeof(T) * (size_t)n * m1, s));

        std::vector<T> H((size_t)m1 * m, T(0)), sn(m, T(0)), cs(m, T(0)), eta(m1, T(0)), hcol((size_t)m1 + 2);

        A(ctx, x, r);
        out.num_matvec++;
        axpby<T>(n, one, b, -one, r, s);
        T r_nrm = std::sqrt(dot<T>(n, r, r, s));
        out.res_norm.push_back((double)r_nrm);
        out.time.push_back(0.0);
        const auto t0 = std::chrono::high_resolution_clock::now();

        if (r_nrm < tol * bnrm) {
            out.success = true;
            if (verbose) {
                std::cout << "After 0 iterations, GMRES achieved rel. residual of " << out.res_norm.back() / bnrm << std::endl;
                std::cout << "GMRES successfully converged within desired tolerance." << std::endl;
            }
            return out;
        }

        Progress bar(maxit);
        if (verbose)
            std::cout << std::setprecision(5) << std::scientific;

        const int rb = red_blocks(n);
        int it = 1;
        for (; it < maxit; ++it) {
            axpby<T>(n, one / r_nrm, r, zero, V, s); // v0 = r / ||r||
            std::fill(eta.begin(), eta.end(), T(0));
            eta[0] = r_nrm;

            int k1 = 0;
            for (int k = 0; k < m; ++k) {
                k1 = k + 1;
                T * vk = V + (size_t)k * n;
                T * w = vk + n;
                A(ctx, vk, w);
                out.num_matvec++;

                // fused MGS sweep: step j computes h_j = <w, v_j> after applying h_{j-1}; the last step applies
                // h_k and produces ||w|| (sqrt taken on the device)
                for (int j = 0; j <= k1; ++j) {
                    const T * vprev = (j > 0) ? V + (size_t)(j - 1) * n : nullptr;
                    const T * vnext = (j < k1) ? V + (size_t)j * n : nullptr;
                    mgs_step_kernel<T><<<rb, RED_THREADS, 0, s>>>(n, w, vprev, vnext, hdev.p + (j > 0 ? j - 1 : 0), wk.part<T>(),
                                                                  wk.ticket.p, hdev.p + j);
                    CB_LAUNCHED();
                }
                scal_inv_dev_kernel<T><<<nblk(n), 256, 0, s>>>(n, hdev.p + k1, w); // w /= ||w|| (inf/nan if 0: caught below)
                CB_LAUNCHED();
                CB_CUDA(cudaMemcpyAsync(hcol.data(), hdev.p, sizeof(T) * (size_t)(k1 + 1), cudaMemcpyDeviceToHost, s));
                CB_CUDA(cudaStreamSynchronize(s));
                T * Hk = &H[(size_t)m1 * k];
                for (int j = 0; j <= k1; ++j)
                    Hk[j] = hcol[j];

                if (Hk[k1] == T(0)) // gmres.cpp:176 (w was all zeros: undo the 0/0 scaling)
                {
                    CB_CUDA(cudaMemsetAsync(w, 0, sizeof(T) * (size_t)n, s));
                    break;
                }

                givens(Hk, cs.data(), sn.data(), k);
                eta[k1] = -sn[k] * eta[k];
                eta[k] = cs[k] * eta[k];

                if (std::abs(eta[k1]) < tol * bnrm)
                    break;
            }

            trsv_upper<T>(k1, H.data(), m1, eta.data());
            CB_CUDA(cudaMemcpyAsync(etadev.p, eta.data(), sizeof(T) * (size_t)k1, cudaMemcpyHostToDevice, s));
            multi_axpy_kernel<T><<<nblk(n), 256, 0, s>>>(n, k1, etadev.p, V, x);
            CB_LAUNCHED();

            A(ctx, x, r);
            out.num_matvec++;
            axpby<T>(n, one, b, -one, r, s);
            r_nrm = std::sqrt(dot<T>(n, r, r, s));
            out.res_norm.push_back((double)r_nrm);
            const auto t1 = std::chrono::high_resolution_clock::now();
            const double dur = 1e-9 * std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
            out.time.push_back(dur);
            if (dur > max_seconds)
                break;

            if (verbose == 1) {
                bar.step();
                std::cout << "[" << bar.bar << "] || iteration " << std::setw(10) << it + 1 << " / " << maxit
                          << " || rel. res. = " << std::setw(10) << r_nrm / bnrm << "\r" << std::flush;
            }
            else if (verbose >= 2)
                std::cout << "iteration " << std::setw(10) << it + 1 << " / " << maxit << " || rel. res. = " << std::setw(10)
                          << r_nrm / bnrm << std::endl;

            if (r_nrm < tol * bnrm) {
                out.success = true;
                break;
            }
        }

        if (verbose == 1)
            std::cout << std::endl;
        if (verbose) {
            std::cout << "After " << it << " iterations, GMRES achieved rel. residual of " << out.res_norm.back() / bnrm << std::endl;
            if (out.success)
                std::cout << "GMRES successfully converged within desired tolerance." << std::endl;
            else
                std::cout << "GMRES failed to converge within desired tolerance." << std::endl;
        }
        out.num_iter = it;
        return out;
    }

    template GmresResult gmres<double>(int64_t, double *, ApplyFn<double>, void *, const double *, int, int, double, int, double, cudaStream_t);
    template GmresResult gmres<float>(int64_t, float *, ApplyFn<float>, void *, const float *, int, int, float, int, double, cudaStream_t);
} // namespace cb200
