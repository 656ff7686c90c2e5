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

// Kernel launches of the training step: programmatic dependent launch.  Every kernel begins with pdl_wait()
// (griddepcontrol.wait: returns once the preceding kernel of the stream has completed and its writes are visible; a
// no-op for a kernel launched without the attribute) and is launched with programmatic stream serialization, so that
// the launch work of kernel k+1 -- and, where the SM has room, its CTAs up to that wait -- overlaps the tail of
// kernel k instead of starting after it.  Inside the step's CUDA graph these become programmatic dependency edges.
// SEDB200_PDL=0 launches plainly (A/B and debugging).
// Measured on the C2 step (profiles/README.md, "r02 programmatic dependent launch"): eagerly launched steps 1.53 ->
// 1.43 ms, but INSIDE a CUDA graph the programmatic edges cost the concurrency of the forked branches (1.40 -> 1.51 ms,
// also with the helper stream launching plainly), while the launch-bound small model gains in both modes (0.61 -> 0.58
// ms as a graph).  So: always on for eager launches; during stream capture only inside a PdlCaptureScope(true), which
// the CRNN entry points open for models without tensor-core conv blocks.
bool pdl_on();
bool& pdl_in_capture();                                      // thread-local
struct PdlCaptureScope {
    bool prev;
    explicit PdlCaptureScope(bool on) : prev(pdl_in_capture()) { pdl_in_capture() = on; }
    ~PdlCaptureScope() { pdl_in_capture() = prev; }
};
// Streams whose launches stay plain: the library's helper stream.  A kernel node with TWO programmatic dependents (the
// next kernel of the caller's stream and the first kernel of a forked branch) loses the concurrency of the branches
// inside a CUDA graph (measured: the side-stream weight gradients stopped overlapping, C2 step 1.40 -> 1.50 ms), so
// forked branches depend on their fork point through ordinary edges.
void pdl_exclude_stream(cudaStream_t st);
bool pdl_excluded(cudaStream_t st);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    bool pdl = pdl_on() && !pdl_excluded(st);
    if (pdl && !pdl_in_capture()) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) pdl = false;
    }
    if (pdl) { cfg.attrs = at; cfg.numAttrs = 1; }
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);      // the error is read by SED_POST_LAUNCH
}
#endif

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
