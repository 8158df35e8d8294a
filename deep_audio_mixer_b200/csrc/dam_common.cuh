// Shared host-side helpers for the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dam_b200.h"

// Thread-local error text behind dam_last_error(); defined in dam_api.cu.
void dam_set_error(const char* fmt, ...);
void dam_clear_error();

#define DAM_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            dam_set_error(__VA_ARGS__);               \
            return DAM_ERR_INVALID_ARGUMENT;          \
        }                                             \
    } while (0)

#define DAM_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            dam_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                      \
            return DAM_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// Every kernel launch of the library is counted (dam_launch_count) so that callers report launches from what
// was dispatched, not from a table.
void dam_note_launch();

// Number of SMs of the current device (cached per device).
int dam_sm_count();

__device__ __forceinline__ float dam_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double dam_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float dam_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double dam_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
