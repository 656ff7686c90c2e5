// head_fused.cu -- the per-frame head of the CRNN as ONE kernel, forward and backward:
//
//   h = relu(d1(x))  ->  logits = d2(h)  ->  p = sigmoid(logits)  ->  frame-wise BCE / focal loss        (forward)
//   dlogits -> dW2, db2, dh (through the ReLU) -> dW1, db1, dx                                            (backward)
//
// Reference: crnn_lightning.py:63-64,72-73 (d1, d2), :27-35 / sed.py:160 (loss), :98 (sigmoid).
// A block owns 128 frames: the GRU output rows, the hidden activations and both weight matrices live in shared
// memory, nothing but x (in), dx / logits / probs (out) and one partial-sum row per block touches HBM.  The
// unfused path needs eight launches and re-reads every intermediate from HBM.  Deterministic: per-block partials,
// fixed-order second pass.
#include "crnn_plan.cuh"
#include "loss_math.cuh"

#include <algorithm>

namespace sedb200 {
namespace {

constexpr int kRows = 128;
constexpr int kThreads = 256;

struct HeadDims { int D, N0, N1, relu; };

__host__ __device__ inline int x_pitch(int D) { return D + 4; }                  // 16 B aligned rows, float4 reads
__host__ __device__ inline size_t head_smem_floats(int D, int N0, int N1) {
    return (size_t)kRows * x_pitch(D) + 2 * (size_t)kRows * (N0 + 1) + 2 * (size_t)kRows * N1 + 2 * (size_t)N0 * D + N0 +
           2 * (size_t)N1 * N0 + N1;
}
__host__ __device__ inline long head_part_floats(int D, int N0, int N1) { return (long)N1 * N0 + N1 + (long)N0 * D + N0 + 1; }

// part row of a block: [dW2 (N1*N0) | db2 (N1) | dW1 (N0*D) | db1 (N0) | loss sum]
// Shared-memory layouts are chosen per phase so that the lanes of a warp read consecutive words or one broadcast
// word: both weight matrices are kept in their own and in transposed order (they are tiny), x rows are 16 B aligned.
__global__ void __launch_bounds__(kThreads)
head_fused_kernel(const float* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ b0,
                  const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ targets,
                  int rows, HeadDims hd, int loss_kind, float alpha, float gamma, float dl_scale,
                  float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ dx,
                  float* __restrict__ part) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int D = hd.D, N0 = hd.N0, N1 = hd.N1, Dp = x_pitch(D), Hp = N0 + 1, D4 = D >> 2;
    float* s_x = sm;                         // [128][D+4]
    float* s_w0 = s_x + kRows * Dp;          // [N0][D]       (16 B aligned: kRows * Dp is a multiple of 4)
    float* s_w0t = s_w0 + N0 * D;            // [D][N0]
    float* s_h = s_w0t + N0 * D;             // [128][N0+1]   relu(d1(x))
    float* s_dh = s_h + kRows * Hp;          // [128][N0+1]
    float* s_dl = s_dh + kRows * Hp;         // [128][N1]
    float* s_t = s_dl + kRows * N1;          // [128][N1]
    float* s_w1 = s_t + kRows * N1;          // [N1][N0]
    float* s_w1t = s_w1 + N1 * N0;           // [N0][N1]
    float* s_b0 = s_w1t + N1 * N0;
    float* s_b1 = s_b0 + N0;
    __shared__ float red[kThreads / 32];
    const int tid = threadIdx.x, r0 = blockIdx.x * kRows, nr = min(kRows, rows - r0);

    for (int i = tid; i < kRows * D4; i += kThreads) {
        const int r = i / D4, k4 = i - r * D4;
        const float4 v = r < nr ? __ldg(reinterpret_cast<const float4*>(x + (long)(r0 + r) * D) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s_x + r * Dp + 4 * k4) = v;
    }
    for (int i = tid; i < kRows * N1; i += kThreads) s_t[i] = (i / N1) < nr ? __ldg(targets + (long)r0 * N1 + i) : 0.0f;
    for (int i = tid; i < N0 * D; i += kThreads) {
        const float w = __ldg(w0 + i);
        const int n = i / D, k = i - n * D;
        s_w0[i] = w;
        s_w0t[k * N0 + n] = w;
    }
    for (int i = tid; i < N1 * N0; i += kThreads) {
        const float w = __ldg(w1 + i);
        const int c = i / N0, n = i - c * N0;
        s_w1[i] = w;
        s_w1t[n * N1 + c] = w;
    }
    if (tid < N0) s_b0[tid] = __ldg(b0 + tid);
    if (tid < N1) s_b1[tid] = __ldg(b1 + tid);
    __syncthreads();

    // ---- h = relu(x W0^T + b0): thread <-> (row, n), n fastest; x row read as float4 (broadcast within a row)
    for (int i = tid; i < kRows * N0; i += kThreads) {
        const int r = i / N0, n = i - r * N0;
        float acc = s_b0[n];
        const float4* xr = reinterpret_cast<const float4*>(s_x + r * Dp);
        for (int k4 = 0; k4 < D4; ++k4) {
            const float4 xv = xr[k4];
            const float* wt = s_w0t + (4 * k4) * N0 + n;
            acc = fmaf(xv.x, wt[0], acc);
            acc = fmaf(xv.y, wt[N0], acc);
            acc = fmaf(xv.z, wt[2 * N0], acc);
            acc = fmaf(xv.w, wt[3 * N0], acc);
        }
        s_h[r * Hp + n] = hd.relu ? fmaxf(acc, 0.0f) : acc;
    }
    __syncthreads();

    // ---- logits, probabilities, loss, dlogits: thread <-> (row, class), class fastest
    float lsum = 0.0f;
    for (int i = tid; i < kRows * N1; i += kThreads) {
        const int r = i / N1, c = i - r * N1;
        float acc = s_b1[c];
        for (int n = 0; n < N0; ++n) acc = fmaf(s_h[r * Hp + n], s_w1t[n * N1 + c], acc);
        float p, loss, dl;
        loss_elem(loss_kind, alpha, gamma, acc, s_t[i], p, loss, dl);
        const bool valid = r < nr;
        s_dl[i] = valid ? dl * dl_scale : 0.0f;
        if (valid) {
            lsum += loss;
            if (logits) logits[(long)r0 * N1 + i] = acc;
            if (probs) probs[(long)r0 * N1 + i] = p;
        }
    }
    lsum = warp_sum(lsum);
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    __syncthreads();

    // ---- dh = (dl W1) through the ReLU: thread <-> (row, n), n fastest
    for (int i = tid; i < kRows * N0; i += kThreads) {
        const int r = i / N0, n = i - r * N0;
        float acc = 0.0f;
        for (int c = 0; c < N1; ++c) acc = fmaf(s_dl[r * N1 + c], s_w1[c * N0 + n], acc);
        if (hd.relu && !(s_h[r * Hp + n] > 0.0f)) acc = 0.0f;
        s_dh[r * Hp + n] = acc;
    }
    __syncthreads();

    // ---- dx = dh W0: thread <-> (row, 4 consecutive k): one broadcast dh word + one float4 of W0 per 4 FMAs
    for (int i = tid; i < kRows * D4; i += kThreads) {
        const int r = i / D4, k4 = i - r * D4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N0; ++n) {
            const float dv = s_dh[r * Hp + n];
            const float4 wv = *reinterpret_cast<const float4*>(s_w0 + n * D + 4 * k4);
            acc.x = fmaf(dv, wv.x, acc.x); acc.y = fmaf(dv, wv.y, acc.y);
            acc.z = fmaf(dv, wv.z, acc.z); acc.w = fmaf(dv, wv.w, acc.w);
        }
        if (r < nr) reinterpret_cast<float4*>(dx + (long)(r0 + r) * D)[k4] = acc;
    }

    // ---- per-block partial weight / bias gradients (rows beyond nr hold zeros in s_dl / s_dh)
    float* pb = part + (long)blockIdx.x * head_part_floats(D, N0, N1);
    for (int i = tid; i < N1 * N0; i += kThreads) {
        const int c = i / N0, n = i - c * N0;
        float acc = 0.0f;
        for (int r = 0; r < kRows; ++r) acc = fmaf(s_dl[r * N1 + c], s_h[r * Hp + n], acc);
        pb[i] = acc;
    }
    pb += N1 * N0;
    for (int c = tid; c < N1; c += kThreads) {
        float acc = 0.0f;
        for (int r = 0; r < kRows; ++r) acc += s_dl[r * N1 + c];
        pb[c] = acc;
    }
    pb += N1;
    // dW1[n][k]: thread <-> (n, 4 consecutive k)
    for (int i = tid; i < N0 * D4; i += kThreads) {
        const int n = i / D4, k4 = i - n * D4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < kRows; ++r) {
            const float dv = s_dh[r * Hp + n];
            const float4 xv = *reinterpret_cast<const float4*>(s_x + r * Dp + 4 * k4);
            acc.x = fmaf(dv, xv.x, acc.x); acc.y = fmaf(dv, xv.y, acc.y);
            acc.z = fmaf(dv, xv.z, acc.z); acc.w = fmaf(dv, xv.w, acc.w);
        }
        float* o = pb + n * D + 4 * k4;                            // part rows are not 16 B aligned in general
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
    }
    pb += N0 * D;
    for (int n = tid; n < N0; n += kThreads) {
        float acc = 0.0f;
        for (int r = 0; r < kRows; ++r) acc += s_dh[r * Hp + n];
        pb[n] = acc;
    }
    if (tid == 0) {
        float t = 0.0f;
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
        pb[N0] = t;
    }
}

// one warp per output: sum the blocks' partials in a fixed order (double accumulator), scatter into the flat
// gradient buffer; the last slot is the loss sum -> mean loss
__global__ void head_reduce_kernel(const float* __restrict__ part, int nblk, HeadDims hd, long n_elems,
                                   float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw0,
                                   float* __restrict__ db0, float* __restrict__ loss) {
    pdl_wait();
    const long stride = head_part_floats(hd.D, hd.N0, hd.N1);
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= stride) return;
    double a = 0.0;
    for (int k = lane; k < nblk; k += 32) a += (double)__ldg(part + (long)k * stride + o);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane != 0) return;
    int i = o;
    if (i < hd.N1 * hd.N0) { dw1[i] = (float)a; return; }
    i -= hd.N1 * hd.N0;
    if (i < hd.N1) { db1[i] = (float)a; return; }
    i -= hd.N1;
    if (i < hd.N0 * hd.D) { dw0[i] = (float)a; return; }
    i -= hd.N0 * hd.D;
    if (i < hd.N0) { db0[i] = (float)a; return; }
    loss[0] = (float)(a / (double)n_elems);
}

}  // namespace

bool head_fused_supported(const Plan& P) {
    if (P.n_dense != 2) return false;
    return head_smem_floats(P.din[0], P.dout[0], P.dout[1]) * 4 <= 200 * 1024 && P.dout[0] <= kThreads &&
           P.dout[1] <= kThreads && P.din[0] % 4 == 0;
}

size_t head_fused_part_floats(const Plan& P, int batch) {
    const long rows = (long)batch * P.T;
    return (size_t)((rows + kRows - 1) / kRows) * head_part_floats(P.din[0], P.dout[0], P.dout[1]);
}

int head_fused_run(const Plan& P, const sedb200_crnn_desc* d, const float* params, int batch, const float* x,
                   const float* targets, int loss_kind, float alpha, float gamma, float grad_scale, float* logits,
                   float* probs, float* loss, float* dx, float* grads, float* part, cudaStream_t st) {
    const HeadDims hd{P.din[0], P.dout[0], P.dout[1], d->dense_relu ? 1 : 0};
    const int rows = batch * P.T;
    const long n_elems = (long)rows * hd.N1;
    const int nblk = (rows + kRows - 1) / kRows;
    const size_t smem = head_smem_floats(hd.D, hd.N0, hd.N1) * 4;
    {
        const int rc = ensure_dyn_smem((const void*)head_fused_kernel, 200 * 1024);
        if (rc) return rc;
    }
    SED_PROF("head.fused", st);
    launch_k(head_fused_kernel, nblk, kThreads, smem, st, x, params + P.dn_w[0], params + P.dn_b[0], params + P.dn_w[1],
                                                   params + P.dn_b[1], targets, rows, hd, loss_kind, alpha, gamma,
                                                   grad_scale / (float)n_elems, logits, probs, dx, part);
    SED_POST_LAUNCH();
    const long outs = head_part_floats(hd.D, hd.N0, hd.N1);
    launch_k(head_reduce_kernel, (int)((outs * 32 + 255) / 256), 256, 0, st, part, nblk, hd, n_elems, grads + P.dn_w[1],
                                                                       grads + P.dn_b[1], grads + P.dn_w[0],
                                                                       grads + P.dn_b[0], loss);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace sedb200
