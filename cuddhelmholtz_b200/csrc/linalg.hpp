// BLAS-1 + GMRES declarations (see linalg.cu).
#pragma once
#include "common.hpp"
#include "dist.hpp"
#include <vector>

namespace cb200
{
    template <typename T> void axpby(int64_t n, T a, const T * x, T b, T * y, cudaStream_t s);
    template <typename T> void scal(int64_t n, T a, T * x, cudaStream_t s);
    template <typename T> void fill(int64_t n, T a, T * x, cudaStream_t s);
    template <typename T> void copy(int64_t n, const T * x, T * y, cudaStream_t s);
    template <typename T> T dot(int64_t n, const T * x, const T * y, cudaStream_t s);   // blocks the host (returns the scalar)
    template <typename T> T dist(int64_t n, const T * x, const T * y, cudaStream_t s);

    // y = A x enqueued on stream s; non-zero return aborts the solve with that status
    template <typename T> using ApplyFn = int (*)(void * ctx, const T * x, T * y, cudaStream_t s);

    struct GmresResult
    {
        bool success;
        int num_iter, num_matvec;
        std::vector<double> res_norm, time;
        // vector traffic of the orthogonalisation (bytes moved by the Gram-Schmidt kernels) and its device time, for bench.py
        double orth_bytes = 0, orth_ms = 0;
        int reorth = 0;     // CGS: second passes taken (cancellation)
        int allreduces = 0; // distributed runs
    };

    enum GmresOrth
    {
        ORTH_MGS = 0, // modified Gram-Schmidt, the reference's arithmetic (source/gmres.cpp:167-172): k+2 fused passes per step
        ORTH_CGS2 = 1 // classical Gram-Schmidt, all inner products in one pass + one update pass; second round on cancellation
    };

    struct GmresOptions
    {
        int orth = ORTH_MGS;
        const Comm * comm = nullptr;             // distributed vectors: inner products are summed over the ranks
        const unsigned char * d_mask = nullptr;  // 1 = this rank owns the entry (counted in inner products), 0 = mirror copy / foreign slot
        bool time_orth = false;                  // record CUDA-event time of the orthogonalisation kernels (adds two event records per step)
        // flexible right preconditioning (FGMRES): z_k = P(v_k) is stored, w = A z_k, x += sum eta_k z_k; P may change from
        // step to step (e.g. an inner iterative solve). ApplyFn<T> cast to void*; null = none.
        void * right_precond = nullptr;
        void * right_precond_ctx = nullptr;
    };

    // Restarted GMRES(m) with the reference's control flow (source/gmres.cpp:91-235). A is a callback that
    // must enqueue y = A x on the stream it is handed.
    template <typename T>
    GmresResult gmres(int64_t n, T * x, ApplyFn<T> A, void * ctx, const T * b, int m, int maxit, T tol, int verbose,
                      double max_seconds, cudaStream_t s, const GmresOptions & opt = GmresOptions());

    // process-wide default of GmresOptions::orth for the entry points without an options argument (set_option "gmres_orth")
    int default_gmres_orth();
    void set_default_gmres_orth(int v);
} // namespace cb200
