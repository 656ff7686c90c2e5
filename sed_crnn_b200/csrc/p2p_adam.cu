// p2p_adam.cu -- the data-parallel exchange step of CRNN training as ONE kernel over NVLink peer memory:
//
//     gradient all-reduce (sum over ranks)  +  global-norm clip  +  Adam            (train_lightning.py:50,
//                                                                                    crnn_lightning.py:195-197)
//
// Every rank's backward pass writes its flat gradient buffer straight into a cudaMalloc'd "exchange region" that the
// other ranks of the node have mapped with CUDA IPC.  The kernel then
//   1. publishes "my gradients for step s are complete" into every peer's flag row (st.release.sys) and waits until
//      all peers have published step s into ITS flag row (ld.acquire.sys, bounded spin),
//   2. reads the gradient slices of all ranks over NVLink (plain P2P loads, rank 0 first -> the sum is bit-identical
//      on every rank, which keeps the replicas' weights identical without a broadcast), stores the sum locally and
//      accumulates the sum of squares,
//   3. grid-syncs (cooperative launch), folds the per-block partials in a fixed order into the global norm,
//   4. applies clip + Adam to its slice of the parameters.
// The gradient buffers are double-buffered by step parity: a rank can only start overwriting buffer b at step s+2
// after its own step s+1 kernel has seen every peer's step-(s+1) flag, which a peer raises only after its step-s kernel
// (the last reader of buffer b) has finished -- so no trailing barrier is needed.
//
// The message is 1.5-9 MB: latency-bound, so each rank simply reads everything (one-shot all-reduce); there is no
// NCCL call, no separate norm / Adam launches and no second pass over the gradients in HBM.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace cg = cooperative_groups;

namespace sedb200 {
namespace {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 256;
constexpr size_t kHeaderBytes = 1024;          // flags[kMaxWorld] (u64) + status word, then the two gradient buffers
constexpr long long kDefaultTimeoutMs = 60 * 1000;   // SEDB200_P2P_TIMEOUT_MS overrides; a peer may legitimately be late
                                                     // (checkpoint on rank 0, validation, a data-loader stall)

struct P2PArgs {
    unsigned char* region[kMaxWorld];   // exchange regions in rank order (own one included), peer-mapped
    int world, rank, parity;
    unsigned long long seq;             // step number published in the flags (monotonic)
    long long timeout_cycles;           // bound of the flag wait
    long n, buf_stride;                 // floats per gradient buffer (padded)
    float* params; float* m; float* v; float* reduced; float* part; float* gnorm;
    float lr, b1, b2, eps, wd, bc1, bc2_sqrt, max_norm, prescale;
    // CUDA-graph steps: exchange number (= optimizer step) and Adam's bias corrections come from the device-side step
    // state (sedb200_step_state: {u64 seed; i64 step; f32 bc1; f32 bc2_sqrt}); null = the by-value fields above
    const unsigned long long* state;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// peer data changes every step: never through the non-coherent path
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kThreads)
p2p_reduce_clip_adam_kernel(const P2PArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sh[kThreads / 32];
    __shared__ float s_coef;
    unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(a.region[a.rank]);
    unsigned int* status = reinterpret_cast<unsigned int*>(a.region[a.rank] + kMaxWorld * 8);
    const unsigned long long seq = a.state ? a.state[1] : a.seq;
    const float bc1 = a.state ? reinterpret_cast<const float*>(a.state)[4] : a.bc1;
    const float bc2_sqrt = a.state ? reinterpret_cast<const float*>(a.state)[5] : a.bc2_sqrt;

    // ---- 1. publish + wait (one thread per peer)
    if (threadIdx.x < a.world) {
        const int p = threadIdx.x;
        if (blockIdx.x == 0) {
            __threadfence_system();     // the backward kernels' gradient stores precede the flag, system-wide
            st_release_sys(reinterpret_cast<unsigned long long*>(a.region[p]) + a.rank, seq);
        }
        const long long t0 = clock64();
        while (ld_acquire_sys(my_flags + p) < seq) {
            if (clock64() - t0 > a.timeout_cycles) {        // a peer died: raise the (sticky) status word -- no hang
                atomicExch(status, 1u);
                __threadfence();
                break;
            }
        }
    }
    __syncthreads();

    // ---- 2. one-shot all-reduce of this block's slices, fixed rank order
    const long n4 = a.n >> 2;                                                   // n is padded to a multiple of 4
    const size_t boff = kHeaderBytes + (size_t)a.parity * a.buf_stride * 4;
    float ss = 0.0f;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long)gridDim.x * kThreads) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < a.world; ++r) {
            const float4 g = ld_peer_f4(reinterpret_cast<const float*>(a.region[r] + boff) + 4 * i);
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
        }
        reinterpret_cast<float4*>(a.reduced)[i] = acc;
        const float x = acc.x * a.prescale, y = acc.y * a.prescale, z = acc.z * a.prescale, w = acc.w * a.prescale;
        ss = fmaf(x, x, ss); ss = fmaf(y, y, ss); ss = fmaf(z, z, ss); ss = fmaf(w, w, ss);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
        a.part[blockIdx.x] = t;
    }
    grid.sync();

    // ---- abort, grid-uniformly, if ANY block of this or an earlier step gave up waiting: the sums above may contain a
    //      peer's stale buffer, so parameters and Adam state stay untouched and the norm reads NaN; the status word is
    //      never cleared, every later step aborts too, and the host raises when it looks (P2PGradExchange.check)
    if (*reinterpret_cast<volatile unsigned int*>(status) != 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.gnorm[0] = __int_as_float(0x7fc00000);
        return;
    }

    // ---- 3. global norm: every block folds the same partials in the same order
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += (double)a.part[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) {
            const float gn = (float)sqrt(s);
            float coef = a.prescale;
            if (a.max_norm > 0.0f) coef *= fminf(a.max_norm / (gn + 1e-6f), 1.0f);
            s_coef = coef;
            if (blockIdx.x == 0) a.gnorm[0] = gn;
        }
    }
    __syncthreads();
    const float coef = s_coef, step_size = a.lr / bc1;

    // ---- 4. Adam (same arithmetic, term for term, as adam_kernel in head_optim.cu)
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long)gridDim.x * kThreads) {
        const float4 g4 = reinterpret_cast<const float4*>(a.reduced)[i];
        float4 p4 = reinterpret_cast<float4*>(a.params)[i];
        float4 m4 = reinterpret_cast<float4*>(a.m)[i];
        float4 v4 = reinterpret_cast<float4*>(a.v)[i];
        const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
        float* ps = &p4.x; float* ms = &m4.x; float* vs = &v4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float w = ps[k];
            float grad = gs[k] * coef;
            grad = fmaf(a.wd, w, grad);
            const float mi = ms[k] + (1.0f - a.b1) * (grad - ms[k]);
            const float vi = fmaf(a.b2, vs[k], (1.0f - a.b2) * grad * grad);
            ms[k] = mi;
            vs[k] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
            ps[k] = w - step_size * (mi / denom);
        }
        reinterpret_cast<float4*>(a.params)[i] = p4;
        reinterpret_cast<float4*>(a.m)[i] = m4;
        reinterpret_cast<float4*>(a.v)[i] = v4;
    }
}

inline long pad4(long n) { return (n + 3) & ~3L; }
inline long buf_stride_floats(long n) { return (pad4(n) + 63) & ~63L; }     // 256 B aligned buffers

}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

size_t sedb200_p2p_region_bytes(long n) { return n <= 0 ? 0 : kHeaderBytes + 2 * (size_t)buf_stride_floats(n) * 4; }

long sedb200_p2p_grad_offset_bytes(long n, int parity) {
    return (long)kHeaderBytes + (long)(parity & 1) * buf_stride_floats(n) * 4;
}

int sedb200_p2p_region_alloc(size_t bytes, void** region_dev, unsigned char* ipc_handle /* [64] */) {
    SED_REQUIRE(bytes >= kHeaderBytes && region_dev && ipc_handle, SEDB200_EINVAL, "p2p_region_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    int rc = require_sm100();
    if (rc) return rc;
    void* p = nullptr;
    SED_CUDA_OK(cudaMalloc(&p, bytes));
    SED_CUDA_OK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    SED_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    std::memcpy(ipc_handle, &h, 64);
    *region_dev = p;
    return SEDB200_OK;
}

int sedb200_p2p_region_open(const unsigned char* ipc_handle, void** region_dev) {
    SED_REQUIRE(ipc_handle && region_dev, SEDB200_EINVAL, "p2p_region_open: bad argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    SED_CUDA_OK(cudaIpcOpenMemHandle(region_dev, h, cudaIpcMemLazyEnablePeerAccess));
    return SEDB200_OK;
}

int sedb200_p2p_region_close(void* region_dev) {
    if (region_dev) SED_CUDA_OK(cudaIpcCloseMemHandle(region_dev));
    return SEDB200_OK;
}

int sedb200_p2p_region_free(void* region_dev) {
    if (region_dev) SED_CUDA_OK(cudaFree(region_dev));
    return SEDB200_OK;
}

int sedb200_p2p_status(const void* region_dev, unsigned int* status_host) {
    SED_REQUIRE(region_dev && status_host, SEDB200_EINVAL, "p2p_status: bad argument");
    SED_CUDA_OK(cudaMemcpy(status_host, reinterpret_cast<const unsigned char*>(region_dev) + kMaxWorld * 8, 4,
                           cudaMemcpyDeviceToHost));
    return SEDB200_OK;
}

long sedb200_p2p_status_offset_bytes(void) { return (long)kMaxWorld * 8; }

size_t sedb200_p2p_scratch_bytes(void) { return 1024 * sizeof(float); }

static int p2p_impl(void* const* regions_host, int world, int rank, long n, long seq, long step, const void* step_state,
                    float* params_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                    float* reduced_dev, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float max_norm, float grad_prescale, float* gnorm_dev,
                    void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(regions_host && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, SEDB200_EINVAL,
                "p2p_allreduce_clip_adam: world=%d rank=%d", world, rank);
    SED_REQUIRE(n % 4 == 0, SEDB200_ESHAPE, "p2p_allreduce_clip_adam: n=%ld is not a multiple of 4 floats", n);
    SED_REQUIRE(n >= 1 && step >= 1 && seq >= 1 && params_dev && exp_avg_dev && exp_avg_sq_dev && reduced_dev && gnorm_dev &&
                    scratch_dev, SEDB200_EINVAL, "p2p_allreduce_clip_adam: bad argument");
    SED_REQUIRE(scratch_bytes >= sedb200_p2p_scratch_bytes(), SEDB200_EWORKSPACE, "p2p_allreduce_clip_adam: scratch");
    int rc = require_sm100();
    if (rc) return rc;
    P2PArgs a;
    std::memset(&a, 0, sizeof(a));
    for (int r = 0; r < world; ++r) {
        SED_REQUIRE(regions_host[r], SEDB200_EINVAL, "p2p_allreduce_clip_adam: region %d is null", r);
        a.region[r] = reinterpret_cast<unsigned char*>(regions_host[r]);
    }
    a.world = world; a.rank = rank; a.parity = (int)(seq & 1); a.seq = (unsigned long long)seq;
    a.state = reinterpret_cast<const unsigned long long*>(step_state);
    a.n = pad4(n); a.buf_stride = buf_stride_floats(n);
    a.params = params_dev; a.m = exp_avg_dev; a.v = exp_avg_sq_dev; a.reduced = reduced_dev;
    a.part = reinterpret_cast<float*>(scratch_dev); a.gnorm = gnorm_dev;
    a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay;
    a.bc1 = (float)(1.0 - std::pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)beta2, (double)step));
    a.max_norm = max_norm; a.prescale = grad_prescale;
    {
        long long ms = kDefaultTimeoutMs;
        if (const char* e = std::getenv("SEDB200_P2P_TIMEOUT_MS")) { const long long v = std::atoll(e); if (v > 0) ms = v; }
        // clock64 ticks at the SM clock; the nominal maximum (2.1 GHz covers every B200 bin) keeps the bound a wall-clock
        // LOWER bound on the wait.  (cudaDevAttrClockRate is a ~1 ms driver query: not something to call per step.)
        a.timeout_cycles = ms * 2100000LL;
    }
    const long n4 = a.n >> 2;
    int grid = (int)std::max<long>(1, std::min<long>((n4 + kThreads - 1) / kThreads, sm_count()));
    grid = std::min(grid, 1024);
    cudaStream_t st = as_stream(stream);
    SED_PROF("p2p_allreduce_clip_adam", st);
    void* kargs[] = {&a};
    SED_CUDA_OK(cudaLaunchCooperativeKernel((void*)p2p_reduce_clip_adam_kernel, dim3(grid), dim3(kThreads), kargs, 0, st));
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_p2p_allreduce_clip_adam(void* const* regions_host, int world, int rank, long n, long seq, long step,
                                    float* params_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                                    float* reduced_dev, float lr, float beta1, float beta2, float eps,
                                    float weight_decay, float max_norm, float grad_prescale, float* gnorm_dev,
                                    void* scratch_dev, size_t scratch_bytes, void* stream) {
    return p2p_impl(regions_host, world, rank, n, seq, step, nullptr, params_dev, exp_avg_dev, exp_avg_sq_dev, reduced_dev,
                    lr, beta1, beta2, eps, weight_decay, max_norm, grad_prescale, gnorm_dev, scratch_dev, scratch_bytes, stream);
}

/* CUDA-graph variant: the exchange number (== optimizer step: one exchange per step since the state was initialised)
 * and the bias corrections are read from the device-side step state; `parity` = (that step) & 1 selects the gradient
 * buffers and is a property of the captured graph (capture one graph per parity and alternate). */
int sedb200_p2p_allreduce_clip_adam_s(void* const* regions_host, int world, int rank, long n, int parity,
                                      const void* step_state_dev, float* params_dev, float* exp_avg_dev,
                                      float* exp_avg_sq_dev, float* reduced_dev, float lr, float beta1, float beta2,
                                      float eps, float weight_decay, float max_norm, float grad_prescale,
                                      float* gnorm_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(step_state_dev && (parity == 0 || parity == 1), SEDB200_EINVAL, "p2p_allreduce_clip_adam_s: bad argument");
    return p2p_impl(regions_host, world, rank, n, 2 + parity, 1, step_state_dev, params_dev, exp_avg_dev, exp_avg_sq_dev,
                    reduced_dev, lr, beta1, beta2, eps, weight_decay, max_norm, grad_prescale, gnorm_dev, scratch_dev,
                    scratch_bytes, stream);
}

}  // extern "C"
