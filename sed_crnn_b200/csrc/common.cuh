// common.cuh -- error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/sedb200.h"

namespace sedb200 {

// thread-local message behind sedb200_last_error()
char* err_buf();
int   fail(int code, const char* fmt, ...);

#define SED_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return ::sedb200::fail(SEDB200_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, \
                                   #expr, cudaGetErrorString(_e));                     \
    } while (0)

#define SED_REQUIRE(cond, code, ...)                                  \
    do {                                                              \
        if (!(cond)) return ::sedb200::fail((code), __VA_ARGS__);     \
    } while (0)

// 0 if the current device is compute capability 10.x, else SEDB200_EARCH (message set).
int  require_sm100();
int  sm_count();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sedb200
