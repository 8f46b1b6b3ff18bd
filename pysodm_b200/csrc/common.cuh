// Shared host-side helpers of the C ABI (error reporting, model registry).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/sodm.h"

namespace sodm {

enum ModelId : int {
    MODEL_SIR_JUMP = 0,
    MODEL_SIR_SI_JUMP = 1,
    MODEL_INFLUENZA_JUMP = 2,
    MODEL_SIR_ODE = 3,
    MODEL_SIR_RATE_ODE = 4,
    MODEL_SIR_SI_ODE = 5,
    MODEL_PPBB_ODE = 6,
    MODEL_PACKED_PFR_ODE = 7,
    MODEL_COUNT = 8
};

char* error_buffer();   // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SODM_CUDA_CHECK(expr)                                                                    \
    do {                                                                                         \
        cudaError_t err__ = (expr);                                                              \
        if (err__ != cudaSuccess)                                                                \
            return ::sodm::fail(SODM_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(err__)); \
    } while (0)

}  // namespace sodm
