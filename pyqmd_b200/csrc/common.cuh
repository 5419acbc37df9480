// common.cuh -- error plumbing shared by the translation units of libpyqmd_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/pyqmd_b200.h"

namespace pyqmd {

void set_error(const char* fmt, ...);

#define PYQMD_CUDA_CHECK(expr)                                                              \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            pyqmd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                      \
            return PYQMD_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)

#define PYQMD_REQUIRE(cond, msg)                                                             \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            pyqmd::set_error("invalid argument: %s (%s)", msg, #cond);                       \
            return PYQMD_ERR_INVALID;                                                        \
        }                                                                                    \
    } while (0)

}  // namespace pyqmd
