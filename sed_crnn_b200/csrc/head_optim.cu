// head_optim.cu -- loss (+gradient) on the logits, clip + Adam over flat buffers, and the integer
// counting behind the segment metrics.
//
// Reference arithmetic:
//   FocalBCELoss            crnn_lightning.py:27-35 (EPS crnn_lightning.py:21)
//   BCEWithLogitsLoss       sed.py:160
//   clip_grad_norm_(1.0)    train_lightning.py:50   (coef = max_norm / (norm + 1e-6), clamped to 1)
//   Adam(lr, wd)            crnn_lightning.py:195-197, sed.py:159 (coupled L2: g += wd * p)
//   threshold + metrics     crnn_lightning.py:112-126, metrics.py:20-68
#include "common.cuh"
#include "loss_math.cuh"
#include <algorithm>
#include <cmath>

namespace sedb200 {
namespace {

constexpr int kRedBlocks = 592;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;                                    // valid on thread 0
}

__global__ void __launch_bounds__(256)
loss_kernel(int kind, float alpha, float gamma, const float* __restrict__ logits, const float* __restrict__ targets,
            long n, float gscale, float* __restrict__ probs, float* __restrict__ dlogits, float* __restrict__ part) {
    pdl_wait();
    __shared__ float sh[8];
    float acc = 0.0f;
    const float inv_n = 1.0f / (float)n;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float l = __ldg(logits + i), t = __ldg(targets + i);
        float p, loss, dl;
        loss_elem(kind, alpha, gamma, l, t, p, loss, dl);
        acc += loss;
        if (probs) probs[i] = p;
        if (dlogits) dlogits[i] = dl * inv_n * gscale;
    }
    const float tot = block_sum_256(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// one warp: fixed lane assignment + fixed shuffle tree (deterministic)
__device__ __forceinline__ double warp_total(const float* __restrict__ part, int nblk) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32) s += (double)part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__global__ void loss_final_kernel(const float* __restrict__ part, int nblk, long n, float* __restrict__ loss) {
    pdl_wait();
    const double s = warp_total(part, nblk);
    if (threadIdx.x == 0) loss[0] = (float)(s / (double)n);
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long n, float prescale, float* __restrict__ part) {
    pdl_wait();
    __shared__ float sh[8];
    float acc = 0.0f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = __ldg(g + i) * prescale;
        acc = fmaf(v, v, acc);
    }
    const float tot = block_sum_256(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// Every block folds the SAME per-block partials of the squared norm in the same order (one warp, fixed lane assignment
// and shuffle tree: the fold p2p_reduce_clip_adam_kernel uses) -- the separate one-warp norm kernel between the two
// passes is gone; block 0 publishes the norm.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float max_norm,
            float prescale, const float* __restrict__ part, int nblk, float* __restrict__ gnorm,
            const float* __restrict__ bc_dev) {
    pdl_wait();
    __shared__ float s_gn;
    if (threadIdx.x < 32) {
        const double s = warp_total(part, nblk);
        if (threadIdx.x == 0) {
            s_gn = (float)sqrt(s);
            if (blockIdx.x == 0) gnorm[0] = s_gn;
        }
    }
    __syncthreads();
    if (bc_dev) { bc1 = __ldg(bc_dev); bc2_sqrt = __ldg(bc_dev + 1); }     // CUDA-graph steps: bias corrections of the device-side step
    float coef = prescale;
    if (max_norm > 0.0f) {
        const float c = max_norm / (s_gn + 1e-6f);
        coef *= fminf(c, 1.0f);
    }
    const float step_size = lr / bc1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float w = p[i];
        float grad = g[i] * coef;
        grad = fmaf(wd, w, grad);
        const float mi = m[i] + (1.0f - b1) * (grad - m[i]);                 // torch: lerp_(grad, 1-beta1)
        const float vi = fmaf(b2, v[i], (1.0f - b2) * grad * grad);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = w - step_size * (mi / denom);
    }
}

// ---- step state on the device (include/sedb200.h: sedb200_step_state): lets a whole training step be captured in a
//      CUDA graph -- the dropout seed and Adam's bias corrections change every step, so they are read from memory.
struct StepState { unsigned long long seed; long long step; float bc1, bc2_sqrt; float pad[2]; };
static_assert(sizeof(StepState) == 32, "sedb200_step_state is 32 bytes");
__global__ void step_state_init_kernel(StepState* s, long long step) {
    pdl_wait();
    s->seed = 0; s->step = step; s->bc1 = 1.0f; s->bc2_sqrt = 1.0f; s->pad[0] = s->pad[1] = 0.0f;
}
__global__ void step_advance_kernel(StepState* s, unsigned long long base_seed, float b1, float b2) {
    pdl_wait();
    const long long step = s->step + 1;
    s->step = step;
    s->seed = base_seed + (unsigned long long)(step - 1);
    s->bc1 = (float)(1.0 - pow((double)b1, (double)step));
    s->bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
}

// ---- threshold + counts.  counts: [0..5] frame TP,Nsys,Nref,S,D,I ; [6..11] block ; [12] block Nref (floor blocks)
__global__ void __launch_bounds__(256)
frame_counts_kernel(const float* __restrict__ probs, const float* __restrict__ targets, long n_rows, int n_cls,
                    float thr, unsigned long long* __restrict__ counts) {
    pdl_wait();
    unsigned long long tp = 0, nsys = 0, nref = 0, S = 0, D = 0, I = 0;
    for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long)gridDim.x * blockDim.x) {
        int fp = 0, fn = 0;
        for (int c = 0; c < n_cls; ++c) {
            const bool o = __ldg(probs + r * n_cls + c) > thr;
            const bool t = __ldg(targets + r * n_cls + c) == 1.0f;
            tp += (o && t); nsys += o; nref += t;
            fp += (o && !t); fn += (!o && t);
        }
        S += min(fp, fn); D += max(0, fn - fp); I += max(0, fp - fn);
    }
    atomicAdd(counts + 0, tp); atomicAdd(counts + 1, nsys); atomicAdd(counts + 2, nref);
    atomicAdd(counts + 3, S); atomicAdd(counts + 4, D); atomicAdd(counts + 5, I);
}

__global__ void __launch_bounds__(256)
block_counts_kernel(const float* __restrict__ probs, const float* __restrict__ targets, long n_rows, int n_cls,
                    int block, float thr, unsigned long long* __restrict__ counts) {
    pdl_wait();
    const long n_ceil = (n_rows + block - 1) / block, n_floor = n_rows / block;
    unsigned long long tp = 0, nsys = 0, nref = 0, S = 0, D = 0, I = 0, nref_er = 0;
    for (long bi = (long)blockIdx.x * blockDim.x + threadIdx.x; bi < n_ceil; bi += (long)gridDim.x * blockDim.x) {
        const long r0 = bi * block, r1 = min(n_rows, r0 + block);
        int fp = 0, fn = 0, nr = 0;
        for (int c = 0; c < n_cls; ++c) {
            bool o = false, t = false;
            for (long r = r0; r < r1; ++r) {
                o |= __ldg(probs + r * n_cls + c) > thr;
                t |= __ldg(targets + r * n_cls + c) == 1.0f;
            }
            tp += (o && t); nsys += o; nref += t; nr += t;
            fp += (o && !t); fn += (!o && t);
        }
        if (bi < n_floor) { S += min(fp, fn); D += max(0, fn - fp); I += max(0, fp - fn); nref_er += nr; }
    }
    atomicAdd(counts + 6, tp); atomicAdd(counts + 7, nsys); atomicAdd(counts + 8, nref);
    atomicAdd(counts + 9, S); atomicAdd(counts + 10, D); atomicAdd(counts + 11, I);
    atomicAdd(counts + 12, nref_er);
}

inline int red_blocks(long n) { return (int)std::max<long>(1, std::min<long>((n + 255) / 256, kRedBlocks)); }

}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

size_t sedb200_loss_scratch_bytes(long n) { (void)n; return (size_t)kRedBlocks * 4; }

int sedb200_loss_fwd_bwd(int kind, float alpha, float gamma, const float* logits, const float* targets, long n,
                         float grad_scale, float* loss, float* probs, float* dlogits, void* scratch,
                         size_t scratch_bytes, void* stream) {
    SED_REQUIRE(kind == SEDB200_LOSS_BCE || kind == SEDB200_LOSS_FOCAL, SEDB200_EINVAL, "loss: kind %d", kind);
    SED_REQUIRE(n >= 1 && logits && targets && loss && scratch, SEDB200_EINVAL, "loss: bad argument");
    SED_REQUIRE(scratch_bytes >= sedb200_loss_scratch_bytes(n), SEDB200_EWORKSPACE, "loss: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int nb = red_blocks(n);
    float* part = reinterpret_cast<float*>(scratch);
    SED_PROF("loss", st);
    launch_k(loss_kernel, nb, 256, 0, st, kind, alpha, gamma, logits, targets, n, grad_scale, probs, dlogits, part);
    SED_POST_LAUNCH();
    launch_k(loss_final_kernel, 1, 32, 0, st, part, nb, n, loss);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

size_t sedb200_clip_adam_scratch_bytes(long n) { (void)n; return (size_t)kRedBlocks * 4; }

static int clip_adam_impl(float* params, const float* grads, float* m, float* v, long n, float lr, float b1, float b2,
                          float eps, float wd, long step, const void* step_state, float max_norm, float prescale,
                          float* gnorm, void* scratch, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(n >= 1 && params && grads && m && v && gnorm && scratch, SEDB200_EINVAL, "clip_adam: bad argument");
    SED_REQUIRE(step >= 1 || step_state, SEDB200_EINVAL, "clip_adam: step %ld (1-based)", step);
    SED_REQUIRE(scratch_bytes >= sedb200_clip_adam_scratch_bytes(n), SEDB200_EWORKSPACE, "clip_adam: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int nb = red_blocks(n);
    float* part = reinterpret_cast<float*>(scratch);
    SED_PROF("clip_adam", st);
    launch_k(sumsq_kernel, nb, 256, 0, st, grads, n, prescale, part);
    SED_POST_LAUNCH();
    const double bc1 = step_state ? 1.0 : 1.0 - std::pow((double)b1, (double)step);
    const double bc2 = step_state ? 1.0 : 1.0 - std::pow((double)b2, (double)step);
    const float* bc_dev = step_state ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(step_state) + 16) : nullptr;
    launch_k(adam_kernel, nb, 256, 0, st, params, grads, m, v, n, lr, b1, b2, eps, wd, (float)bc1, (float)std::sqrt(bc2),
                                    max_norm, prescale, part, nb, gnorm, bc_dev);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_clip_adam(float* params, const float* grads, float* m, float* v, long n, float lr, float b1, float b2,
                      float eps, float wd, long step, float max_norm, float prescale, float* gnorm, void* scratch,
                      size_t scratch_bytes, void* stream) {
    return clip_adam_impl(params, grads, m, v, n, lr, b1, b2, eps, wd, step, nullptr, max_norm, prescale, gnorm, scratch,
                          scratch_bytes, stream);
}
int sedb200_clip_adam_s(float* params, const float* grads, float* m, float* v, long n, float lr, float b1, float b2,
                        float eps, float wd, const void* step_state, float max_norm, float prescale, float* gnorm,
                        void* scratch, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(step_state, SEDB200_EINVAL, "clip_adam_s: null step state");
    return clip_adam_impl(params, grads, m, v, n, lr, b1, b2, eps, wd, 0, step_state, max_norm, prescale, gnorm, scratch,
                          scratch_bytes, stream);
}

size_t sedb200_step_state_bytes(void) { return sizeof(StepState); }
int sedb200_step_state_init(void* step_state, long step, void* stream) {
    SED_REQUIRE(step_state && step >= 0, SEDB200_EINVAL, "step_state_init: bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    launch_k(step_state_init_kernel, 1, 1, 0, as_stream(stream), reinterpret_cast<StepState*>(step_state), step);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}
int sedb200_step_advance(void* step_state, unsigned long long base_seed, float b1, float b2, void* stream) {
    SED_REQUIRE(step_state, SEDB200_EINVAL, "step_advance: null step state");
    int rc = require_sm100();
    if (rc) return rc;
    launch_k(step_advance_kernel, 1, 1, 0, as_stream(stream), reinterpret_cast<StepState*>(step_state), base_seed, b1, b2);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_threshold_counts(const float* probs, const float* targets, long n_rows, int n_cls, int block,
                             float threshold, unsigned long long* counts, void* stream) {
    SED_REQUIRE(n_rows >= 0 && n_cls >= 1 && block >= 1 && counts, SEDB200_EINVAL, "threshold_counts: bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    SED_CUDA_OK(cudaMemsetAsync(counts, 0, 13 * sizeof(unsigned long long), st));
    if (n_rows == 0) return SEDB200_OK;
    SED_REQUIRE(probs && targets, SEDB200_EINVAL, "threshold_counts: null buffer");
    launch_k(frame_counts_kernel, red_blocks(n_rows), 256, 0, st, probs, targets, n_rows, n_cls, threshold, counts);
    SED_POST_LAUNCH();
    const long nblk = (n_rows + block - 1) / block;
    launch_k(block_counts_kernel, red_blocks(nblk), 256, 0, st, probs, targets, n_rows, n_cls, block, threshold, counts);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // extern "C"
