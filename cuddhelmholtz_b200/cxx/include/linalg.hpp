// BLAS-1 front end (reference include/linalg.hpp:16-54): device pointers in, scalars out on the host. Thin inline
// wrappers over the C ABI; dot/dist are deterministic two-stage reductions in the library.
#ifndef CUDDH_LINALG_HPP
#define CUDDH_LINALG_HPP

#include <assert.h>
#include <cmath>

#include <cuda_runtime.h>

#include "cuddh_config.hpp"
#include "cuddh_error.hpp"
#include "forall.hpp"
#include "HostDeviceArray.hpp"

namespace cuddh
{
    /// y[i] <- a * x[i] + b * y[i]
    inline void axpby(int n, double a, const double * x, double b, double * y) { cuddh_check(cuddh_b200_axpby_d(n, a, x, b, y, nullptr)); }
    inline void axpby(int n, float a, const float * x, float b, float * y) { cuddh_check(cuddh_b200_axpby_f(n, a, x, b, y, nullptr)); }

    inline double dot(int n, const double * x, const double * y) { double r = 0; cuddh_check(cuddh_b200_dot_d(n, x, y, &r, nullptr)); return r; }
    inline float dot(int n, const float * x, const float * y) { float r = 0; cuddh_check(cuddh_b200_dot_f(n, x, y, &r, nullptr)); return r; }

    inline double norm(int n, const double * x) { return std::sqrt(dot(n, x, x)); }
    inline float norm(int n, const float * x) { return std::sqrt(dot(n, x, x)); }

    /// ||x - y||
    inline double dist(int n, const double * x, const double * y) { double r = 0; cuddh_check(cuddh_b200_dist_d(n, x, y, &r, nullptr)); return r; }
    inline float dist(int n, const float * x, const float * y) { float r = 0; cuddh_check(cuddh_b200_dist_f(n, x, y, &r, nullptr)); return r; }

    inline void copy(int n, const double * x, double * y) { cuddh_check(cuddh_b200_copy_d(n, x, y, nullptr)); }
    inline void copy(int n, const float * x, float * y) { cuddh_check(cuddh_b200_copy_f(n, x, y, nullptr)); }
    inline void copy(int n, const int * x, int * y) { cuddh_check(cuddh_b200_copy_i(n, x, y, nullptr)); }

    inline void scal(int n, double a, double * x) { cuddh_check(cuddh_b200_scal_d(n, a, x, nullptr)); }
    inline void scal(int n, float a, float * x) { cuddh_check(cuddh_b200_scal_f(n, a, x, nullptr)); }

    inline void fill(int n, double a, double * x) { cuddh_check(cuddh_b200_fill_d(n, a, x, nullptr)); }
    inline void fill(int n, float a, float * x) { cuddh_check(cuddh_b200_fill_f(n, a, x, nullptr)); }
    inline void fill(int n, int a, int * x) { cuddh_check(cuddh_b200_fill_i(n, a, x, nullptr)); }

    inline void zeros(int n, double * x) { cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), 0); }
    inline void zeros(int n, float * x) { cudaMemsetAsync(x, 0, (size_t)n * sizeof(float), 0); }
    inline void zeros(int n, int * x) { cudaMemsetAsync(x, 0, (size_t)n * sizeof(int), 0); }

    inline void ones(int n, double * x) { fill(n, 1.0, x); }
    inline void ones(int n, float * x) { fill(n, 1.0f, x); }
    inline void ones(int n, int * x) { cudaMemset(x, 1, (size_t)n * sizeof(int)); } // byte pattern 0x01, as the reference (include/linalg.hpp:54)
} // namespace cuddh

#endif
