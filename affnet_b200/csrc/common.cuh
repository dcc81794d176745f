// Shared helpers for the affnet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/affnet_b200.h"

namespace ag {

void set_error(const char* fmt, ...);
extern thread_local int g_launches;  // kernels launched since last reset (host-side counter)

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return AG_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
               ? AG_ERR_NO_DEVICE
               : AG_ERR_CUDA;
}

void prof_mark(const char* name);  // records an event after a launch when profiling is on (see ag_prof_begin)

#define AG_CHECK_LAUNCH(name)                                        \
    do {                                                             \
        ::ag::g_launches++;                                          \
        int _rc = ::ag::check_cuda(cudaGetLastError(), name);        \
        if (_rc != AG_OK) return _rc;                                \
        ::ag::prof_mark(name);                                       \
    } while (0)

#define AG_REQUIRE(cond, msg)                                        \
    do {                                                             \
        if (!(cond)) {                                               \
            ::ag::set_error("%s: %s", __func__, msg);                \
            return AG_ERR_INVALID;                                   \
        }                                                            \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Order-preserving float -> uint32 (larger float => larger uint), total order incl. negatives.
__host__ __device__ inline uint32_t float_to_ordered(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace ag
