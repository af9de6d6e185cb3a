// cuddh_error: print a banner and abort (reference source/cuddh_error.cpp:5-9); also the bridge from C-ABI status
// codes to that behaviour.
#ifndef CUDDH_ERROR_HPP
#define CUDDH_ERROR_HPP

#include <assert.h>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "cuddh_config.hpp"
#include "cuddh_b200.h"

namespace cuddh
{
    __host__ __device__ inline void cuddh_error(const char * msg)
    {
        printf("--- CUDDH ERROR ---\n\t%s\n-------------------\n", msg);
        assert(0);
    }

    // non-zero C-ABI status -> cuddh_error with the library's message
    inline void cuddh_check(int status)
    {
        if (status != 0)
            cuddh_error(cuddh_b200_last_error());
    }
} // namespace cuddh

#endif
