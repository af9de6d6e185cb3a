// BLAS-1 + GMRES declarations (see linalg.cu).
#pragma once
#include "common.hpp"
#include <vector>

namespace cb200
{
    template <typename T> void axpby(int64_t n, T a, const T * x, T b, T * y, cudaStream_t s);
    template <typename T> void scal(int64_t n, T a, T * x, cudaStream_t s);
    template <typename T> void fill(int64_t n, T a, T * x, cudaStream_t s);
    template <typename T> void copy(int64_t n, const T * x, T * y, cudaStream_t s);
    template <typename T> T dot(int64_t n, const T * x, const T * y, cudaStream_t s);   // blocks the host (returns the scalar)
    template <typename T> T dist(int64_t n, const T * x, const T * y, cudaStream_t s);

    template <typename T> using ApplyFn = void (*)(void * ctx, const T * x, T * y);

    struct GmresResult
    {
        bool success;
        int num_iter, num_matvec;
        std::vector<double> res_norm, time;
    };

    // Restarted GMRES(m) with the reference's control flow (source/gmres.cpp:91-235). A is a callback that
    // must enqueue y = A x on stream s (or on a stream ordered with it, e.g. the legacy default stream).
    template <typename T>
    GmresResult gmres(int64_t n, T * x, ApplyFn<T> A, void * ctx, const T * b, int m, int maxit, T tol, int verbose,
                      double max_seconds, cudaStream_t s);
} // namespace cb200
