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

// every kernel launch goes through this: counts the launch and surfaces launch errors
void count_launch();
#define SED_POST_LAUNCH()                        \
    do {                                         \
        ::sedb200::count_launch();               \
        SED_CUDA_OK(cudaGetLastError());         \
    } while (0)

// optional phase profiler (sedb200_prof_enable): CUDA events on the launching stream around a phase
bool prof_on();
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(const char* name, cudaStream_t s) : st(s), on(prof_on()) { if (on) prof_begin(name, st); }
    ~ProfScope() { if (on) prof_end(st); }
};
#define SED_PROF(name_literal, stream) ::sedb200::ProfScope _sed_prof(name_literal, stream)

// 0 if the current device is compute capability 10.x, else SEDB200_EARCH (message set).
int  require_sm100();
int  sm_count();
// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes) once per (device, kernel): the attribute belongs to
// the device's context, so a process that drives several GPUs must set it on each of them.  Thread-safe.
int  ensure_dyn_smem(const void* func, int bytes);

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------ device helpers
// p[0] + p[stride] + ... (n terms) added in ascending order, U loads in flight at a time: the partial-sum
// finalizers are chains of dependent global loads otherwise (one L2 round trip per term)
template <int U, typename Acc>
__device__ __forceinline__ Acc ordered_sum(const float* __restrict__ p, long stride, int n) {
    Acc a = (Acc)0;
    int k = 0;
    for (; k + U <= n; k += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(p + (long)(k + u) * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) a += (Acc)v[u];
    }
    for (; k < n; ++k) a += (Acc)__ldg(p + (long)k * stride);
    return a;
}

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
