// Shared helpers for libnkbk.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/nkbk.h"

namespace nkbk {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool k1_overlap_previous();
bool heads_one_launch();      // nkbk_heads_one_launch(): may the heads calls use the one-launch persistent kernel?
//   // nkbk_k1_overlap_previous(): K1 launches as programmatic dependents

#define NKBK_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            nkbk::set_error(__VA_ARGS__);    \
            return NKBK_E_ARG;               \
        }                                    \
    } while (0)

#define NKBK_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            nkbk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
            return NKBK_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// Launch check: catches configuration errors at enqueue time without synchronising.
#define NKBK_CHECK_LAUNCH(name)                                                            \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            nkbk::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return NKBK_E_CUDA;                                                            \
        }                                                                                  \
        nkbk::count_launch();                                                              \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}

}  // namespace nkbk
