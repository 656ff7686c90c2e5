// crnn.cu -- CRNN forward / backward orchestration and the conv-block elementwise kernels.
//
// Reference arithmetic (see include/sedb200.h): crnn_lightning.py:41-73, sed.py:82-112.
// Internal activation layout is channels-last [B][H][W][C]; the last conv block writes straight
// into the GRU-ready [B][T][flat] layout (feature = c*F + f), which removes the reference's
// permute+reshape copy (crnn_lightning.py:68-70).
//
// This file holds the fp32-exact path: contractions run through the functor GEMM of gemm_simt.cuh.
#include <cuda_bf16.h>
#include "crnn_plan.cuh"
#include <map>
#include <mutex>
#include "gemm_simt.cuh"
#include "conv_small.cuh"
#include "gru_scan.cuh"
#include "tc_conv.cuh"
#include "tc_gemm.cuh"
#include "crnn_block.cuh"
#include "conv0_lean.cuh"

#include <algorithm>
#include <cstdlib>

namespace sedb200 {
namespace {

// ----------------------------------------------------------------------------- GEMM operand functors
struct RowMajor {                       // op(row, k) = p[row*ld + k]
    static constexpr bool kContigK = true;
    const float* p; long ld;
    __device__ float operator()(int r, int k) const { return __ldg(p + (long)r * ld + k); }
};
struct ColMajor {                       // op(row, k) = p[k*ld + row]
    static constexpr bool kContigK = false;
    const float* p; long ld;
    __device__ float operator()(int r, int k) const { return __ldg(p + (long)k * ld + r); }
};

// im2col view of a conv input with arbitrary strides (NCHW user input or channels-last activation):
// A(m, k), m = (b,h,w), k = tap*Cin + ci
struct ConvFwdA {
    static constexpr bool kContigK = true;
    const float* in; int H, W, Cin; long sB, sH, sW, sC;
    __device__ float operator()(int m, int k) const {
        const int w = m % W, t = m / W, h = t % H, b = t / H;
        const int ci = k % Cin, tap = k / Cin, r = tap / 3, s = tap - 3 * r;
        const int hh = h + r - 1, ww = w + s - 1;
        if ((unsigned)hh >= (unsigned)H || (unsigned)ww >= (unsigned)W) return 0.0f;
        return __ldg(in + b * sB + hh * sH + ww * sW + ci * sC);
    }
};
struct ConvFwdB {                       // B(n, k) = weight[n][ci][tap]
    static constexpr bool kContigK = false;
    const float* w; int Cin;
    __device__ float operator()(int n, int k) const {
        const int ci = k % Cin, tap = k / Cin;
        return __ldg(w + ((long)n * Cin + ci) * 9 + tap);
    }
};
// dgrad: dIn(m=(b,h,w), ci) = sum_{tap,co} dY[b][h-r+1][w-s+1][co] * weight[co][ci][tap]
struct ConvDgradA {                     // k = tap*C + co
    static constexpr bool kContigK = true;
    const float* dy; int H, W, C;
    __device__ float operator()(int m, int k) const {
        const int w = m % W, t = m / W, h = t % H, b = t / H;
        const int co = k % C, tap = k / C, r = tap / 3, s = tap - 3 * r;
        const int hh = h - r + 1, ww = w - s + 1;
        if ((unsigned)hh >= (unsigned)H || (unsigned)ww >= (unsigned)W) return 0.0f;
        return __ldg(dy + (((long)b * H + hh) * W + ww) * C + co);
    }
};
struct ConvDgradB {                     // B(n=ci, k=(tap,co))
    static constexpr bool kContigK = false;
    const float* w; int Cin, C;
    __device__ float operator()(int n, int k) const {
        const int co = k % C, tap = k / C;
        return __ldg(w + ((long)co * Cin + n) * 9 + tap);
    }
};
// wgrad: dW(co, j=ci*9+tap) = sum_p dY[p][co] * In[p shifted by tap][ci]
struct ConvWgradB {                     // B(n=j, k=p)
    static constexpr bool kContigK = false;
    const float* in; int H, W; long sB, sH, sW, sC;
    __device__ float operator()(int n, int k) const {
        const int ci = n / 9, tap = n - 9 * ci, r = tap / 3, s = tap - 3 * r;
        const int w = k % W, t = k / W, h = t % H, b = t / H;
        const int hh = h + r - 1, ww = w + s - 1;
        if ((unsigned)hh >= (unsigned)H || (unsigned)ww >= (unsigned)W) return 0.0f;
        return __ldg(in + b * sB + hh * sH + ww * sW + ci * sC);
    }
};
// h_{t-1} of one GRU direction, as B(n=j, k=(b,t)) for dW_hh
struct HPrevB {
    static constexpr bool kContigK = false;
    const float* out; int T, H, dir;
    __device__ float operator()(int n, int k) const {
        const int t = k % T;
        const int tp = dir ? t + 1 : t - 1;
        if (tp < 0 || tp >= T) return 0.0f;
        return __ldg(out + ((long)(k - t + tp)) * 2 * H + dir * H + n);
    }
};

struct EpiStore {                       // out[m*ld + n] = act(acc + bias[n])
    float* out; long ld; const float* bias; int relu;
    __device__ void operator()(int m, int n, float acc, int) const {
        float v = acc + (bias ? __ldg(bias + n) : 0.0f);
        if (relu) v = fmaxf(v, 0.0f);
        out[(long)m * ld + n] = v;
    }
};
struct EpiPartial {                     // part[z][m*ld + n] = acc
    float* part; long mn, ld;
    __device__ void operator()(int m, int n, float acc, int z) const { part[(long)z * mn + (long)m * ld + n] = acc; }
};
struct EpiReluMask {                    // out = act > 0 ? acc : 0   (gradient through a ReLU'd dense layer)
    float* out; long ld; const float* act;
    __device__ void operator()(int m, int n, float acc, int) const {
        const long i = (long)m * ld + n;
        out[i] = __ldg(act + i) > 0.0f ? acc : 0.0f;
    }
};
struct EpiStrided {                     // dx in NCHW from m = (b,h,w), n = ci
    float* out; int H, W; long sB, sH, sW, sC;
    __device__ void operator()(int m, int n, float acc, int) const {
        const int w = m % W, t = m / W, h = t % H, b = t / H;
        out[b * sB + h * sH + w * sW + n * sC] = acc;
    }
};

// ----------------------------------------------------------------------------- BatchNorm statistics
// stat layout: [0,C) mean, [C,2C) invstd, [2C,3C) scale = gamma*invstd, [3C,4C) shift = beta - mean*scale
// deterministic block-per-channel reduction of the per-block partials [nblk][2][C]: fixed thread assignment,
// fixed shuffle tree, fixed cross-warp order => bit-reproducible.  blockDim.x == 128.
__device__ __forceinline__ void reduce_pair(const float* __restrict__ part, int nblk, int C, int c, double& s,
                                            double& ss) {
    __shared__ double sh[2][4];
    // thread t owns partials t, t + 128, ...
    const int mine = nblk > (int)threadIdx.x ? (nblk - (int)threadIdx.x + 127) / 128 : 0;
    double a = ordered_sum<8, double>(part + ((long)threadIdx.x * 2 + 0) * C + c, 256L * C, mine);
    double b = ordered_sum<8, double>(part + ((long)threadIdx.x * 2 + 1) * C + c, 256L * C, mine);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    s = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
    ss = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
}

// one 128-thread block per channel
__global__ void __launch_bounds__(128)
bn_finalize_train_kernel(const float* __restrict__ part, int nblk, int C, long n,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float eps, float momentum, float* __restrict__ running,
                                         float* __restrict__ stat) {
    pdl_wait();
    const int c = blockIdx.x;
    double s, ss;
    reduce_pair(part, nblk, C, c, s, ss);
    if (threadIdx.x != 0) return;
    const double mean = s / (double)n;
    double var = ss / (double)n - mean * mean;          // biased (normalisation)
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    stat[c] = (float)mean;
    stat[C + c] = invstd;
    stat[2 * C + c] = sc;
    stat[3 * C + c] = beta[c] - (float)mean * sc;
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running[c] = (1.0f - momentum) * running[c] + momentum * (float)mean;
    running[C + c] = (1.0f - momentum) * running[C + c] + momentum * (float)unbiased;
}

__global__ void bn_finalize_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                        float eps, const float* __restrict__ running, float* __restrict__ stat) {
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float mean = running[c];
    const float invstd = 1.0f / sqrtf(running[C + c] + eps);
    const float sc = gamma[c] * invstd;
    stat[c] = mean;
    stat[C + c] = invstd;
    stat[2 * C + c] = sc;
    stat[3 * C + c] = beta[c] - mean * sc;
}

// ----------------------------------------------------------------------------- BN + ReLU + max-pool(1,p) (+dropout)
__global__ void __launch_bounds__(256)
bn_relu_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ stat, float* __restrict__ out,
                        __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long n_vec, PoolGeom g) {
    pdl_wait();
    const int C4 = g.C >> 2;
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long t = i / C4;
        const int wo = (int)(t % g.Wo); t /= g.Wo;
        const int h = (int)(t % g.H);
        const long b = t / g.H;
        const int c = c4 * 4;
        const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
        const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
        const float* src = y + (((b * g.H + h) * g.W) + (long)wo * g.p) * g.C + c;
        float4 m = make_float4(0.0f, 0.0f, 0.0f, 0.0f);           // relu floor
        for (int j = 0; j < g.p; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long)j * g.C));
            m.x = fmaxf(m.x, fmaf(v.x, sc.x, sh.x));
            m.y = fmaxf(m.y, fmaf(v.y, sc.y, sh.y));
            m.z = fmaxf(m.z, fmaf(v.z, sc.z, sh.z));
            m.w = fmaxf(m.w, fmaf(v.w, sc.w, sh.w));
        }
        if (g.drop_p > 0.0f) {
            const Keep4 kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
            m.x = kp.k[0] ? m.x * keep_scale : 0.0f;
            m.y = kp.k[1] ? m.y * keep_scale : 0.0f;
            m.z = kp.k[2] ? m.z * keep_scale : 0.0f;
            m.w = kp.k[3] ? m.w * keep_scale : 0.0f;
        }
        if (out_hi) store_planes4(out_hi, out_lo, i / C4, c4, C4, m);   // channels-last planes
        if (out) {
            float* dst = out + b * g.oB + h * g.oH + wo * g.oW + c * g.oC;
            if (g.oC == 1) {
                *reinterpret_cast<float4*>(dst) = m;
            } else {
                dst[0] = m.x; dst[g.oC] = m.y; dst[2 * g.oC] = m.z; dst[3 * g.oC] = m.w;
            }
        }
    }
}

// Same kernel with a compile-time pool width and 32-bit index arithmetic (the P loads of a window are all in flight
// before the first max; no 64-bit divisions per element).
template <int P>
__global__ void __launch_bounds__(256)
bn_relu_pool_fwd_t_kernel(const float* __restrict__ y, const float* __restrict__ stat, float* __restrict__ out,
                          __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, unsigned n_pix,
                          PoolGeom g) {
    pdl_wait();
    const int C4 = g.C >> 2, rows = 256 / C4;
    const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4, c = c4 * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    auto fetch = [&](unsigned pix, float4 (&v)[P]) {
        unsigned b, h, wo;
        pix_bhw(pix, g, b, h, wo);
#pragma unroll
        for (int j = 0; j < P; ++j)
            v[j] = __ldg(reinterpret_cast<const float4*>(y + ((((long)b * g.H + h) * g.W) + (long)wo * P + j) * g.C + c));
    };
    auto finish = [&](unsigned pix, const float4 (&v)[P]) {
        unsigned b, h, wo;
        pix_bhw(pix, g, b, h, wo);
        float4 m = make_float4(0.0f, 0.0f, 0.0f, 0.0f);           // relu floor
#pragma unroll
        for (int j = 0; j < P; ++j) {
            m.x = fmaxf(m.x, fmaf(v[j].x, sc.x, sh.x));
            m.y = fmaxf(m.y, fmaf(v[j].y, sc.y, sh.y));
            m.z = fmaxf(m.z, fmaf(v[j].z, sc.z, sh.z));
            m.w = fmaxf(m.w, fmaf(v[j].w, sc.w, sh.w));
        }
        const long i = (long)pix * C4 + c4;                        // same element numbering as the generic kernel
        if (g.drop_p > 0.0f) {
            const Keep4 kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
            m.x = kp.k[0] ? m.x * keep_scale : 0.0f;
            m.y = kp.k[1] ? m.y * keep_scale : 0.0f;
            m.z = kp.k[2] ? m.z * keep_scale : 0.0f;
            m.w = kp.k[3] ? m.w * keep_scale : 0.0f;
        }
        if (out_hi) store_planes4(out_hi, out_lo, (long)pix, c4, C4, m);
        if (out) {
            float* dst = out + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC;
            if (g.oC == 1) {
                *reinterpret_cast<float4*>(dst) = m;
            } else {
                dst[0] = m.x; dst[g.oC] = m.y; dst[2 * g.oC] = m.z; dst[3 * g.oC] = m.w;
            }
        }
    };
    // two windows per trip: all 2 P loads are in flight before the first use
    const unsigned stride = gridDim.x * rows;
    unsigned pix = blockIdx.x * rows + prow;
    for (; pix + stride < n_pix; pix += 2 * stride) {
        float4 v0[P], v1[P];
        fetch(pix, v0);
        fetch(pix + stride, v1);
        finish(pix, v0);
        finish(pix + stride, v1);
    }
    if (pix < n_pix) {
        float4 v0[P];
        fetch(pix, v0);
        finish(pix, v0);
    }
}

// shared by the two backward passes: for one pooling window and 4 channels, recompute the winner and
// return dz at the winner (0 elsewhere) plus the winner index.
struct WindowGrad { float dz[4]; int arg[4]; };

__device__ __forceinline__ WindowGrad window_grad(const float* __restrict__ src, const float* __restrict__ dA,
                                                  const float4 sc, const float4 sh, const PoolGeom& g, long i) {
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    WindowGrad r;
#pragma unroll
    for (int q = 0; q < 4; ++q) r.arg[q] = 0;
    for (int j = 0; j < g.p; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long)j * g.C));
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float z = fmaf(vv[q], scv[q], shv[q]);
            if (z > best[q]) { best[q] = z; r.arg[q] = j; }      // first maximum wins (PyTorch max_pool2d)
        }
    }
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    Keep4 kp;
    if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float gq = __ldg(dA + q * g.oC);
        if (g.drop_p > 0.0f) gq = kp.k[q] ? gq * keep_scale : 0.0f;
        r.dz[q] = best[q] > 0.0f ? gq : 0.0f;                       // ReLU gate
    }
    return r;
}


// Compile-time pool width: the P conv outputs of a window are loaded first (P independent 16 B loads in
// flight per thread), then evaluated from registers -- no second trip to memory for the winner.
template <int P>
__device__ __forceinline__ void load_window(const float* __restrict__ src, int C, float4 (&v)[P]) {
#pragma unroll
    for (int j = 0; j < P; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(src + (long)j * C));
}
template <int P>
__device__ __forceinline__ void eval_window(const float4 (&v)[P], const float (&scv)[4], const float (&shv)[4],
                                            float (&gq)[4], const PoolGeom& g, long i, float (&dz)[4], int (&arg)[4],
                                            float (&yarg)[4]) {
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int q = 0; q < 4; ++q) { arg[q] = 0; yarg[q] = 0.0f; }
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float z = fmaf(vv[q], scv[q], shv[q]);
            if (z > best[q]) { best[q] = z; arg[q] = j; yarg[q] = vv[q]; }
        }
    }
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    Keep4 kp;
    if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float t = gq[q];
        if (g.drop_p > 0.0f) t = kp.k[q] ? t * keep_scale : 0.0f;
        dz[q] = best[q] > 0.0f ? t : 0.0f;
    }
}

template <int P>
__global__ void __launch_bounds__(256)
bn_pool_bwd_sums_t_kernel(const float* __restrict__ y, const float* __restrict__ stat, const float* __restrict__ dA,
                          long n_pix, PoolGeom g, float* __restrict__ part) {
    pdl_wait();
    __shared__ float4 s1[256], s2[256];
    const int C4 = g.C >> 2, rows = 256 / C4;
    const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4;
    const int c = c4 * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
    float a[4] = {0, 0, 0, 0}, bsum[4] = {0, 0, 0, 0};
    const unsigned Wo = (unsigned)g.Wo, H = (unsigned)g.H, npx = (unsigned)n_pix;     // n_pix < 2^31 (launcher)
    for (unsigned pix = blockIdx.x * rows + prow; pix < npx; pix += gridDim.x * rows) {
        unsigned b, h, wo;
        pix_bhw(pix, g, b, h, wo);
        float4 v[P];
        float gq[4], dz[4], yarg[4];
        int arg[4];
        load_window<P>(y + ((((long)b * g.H + h) * g.W) + (long)wo * P) * g.C + c, g.C, v);
        load_dA(dA + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC, g.oC, gq);
        eval_window<P>(v, scv, shv, gq, g, (long)pix * C4 + c4, dz, arg, yarg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] += dz[q];
            bsum[q] = fmaf(dz[q], (yarg[q] - muv[q]) * isv[q], bsum[q]);
        }
    }
    s1[threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
    s2[threadIdx.x] = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
    __syncthreads();
    if (prow == 0) {
        float4 ta = make_float4(0, 0, 0, 0), tb = make_float4(0, 0, 0, 0);
        for (int r = 0; r < rows; ++r) {
            const float4 u = s1[r * C4 + c4], w = s2[r * C4 + c4];
            ta.x += u.x; ta.y += u.y; ta.z += u.z; ta.w += u.w;
            tb.x += w.x; tb.y += w.y; tb.z += w.z; tb.w += w.w;
        }
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 0) * g.C + c) = ta;
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 1) * g.C + c) = tb;
    }
}

// pass 1 from the block OUTPUT instead of the conv output: the saved activation a = dropout(relu(max_window(bn(y))))
// already says everything the two sums need -- a > 0 <=> the window's winner passed the ReLU and was kept by the
// dropout, in which case dz = dA * keep_scale and xhat(winner) = (a / keep_scale - beta) / gamma.  Reads
// |a| + |dA| (1/p-th of the conv output each) instead of the whole conv output, and needs neither the window
// search nor the dropout generator.  `act` (fp32, block-output strides) or its bf16 hi/lo planes (channels-last).
__global__ void __launch_bounds__(256)
bn_bwd_sums_act_kernel(const float* __restrict__ act, const __nv_bfloat16* __restrict__ act_hi,
                       const __nv_bfloat16* __restrict__ act_lo, const float* __restrict__ stat,
                       const float* __restrict__ dA, unsigned n_pix, PoolGeom g, float* __restrict__ part,
                       float* __restrict__ amax_part) {
    pdl_wait();
    __shared__ float4 s1[256], s2[256];
    const int C4 = g.C >> 2, rows = 256 / C4;
    const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4, c = c4 * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    // xhat = (z - sh) * (invstd / sc) - mean * invstd   (z = sc * y + sh);  sc == 0 (gamma == 0) carries no information
    const float shv[4] = {sh.x, sh.y, sh.z, sh.w};
    const float rs[4] = {sc.x != 0.f ? is.x / sc.x : 0.f, sc.y != 0.f ? is.y / sc.y : 0.f,
                         sc.z != 0.f ? is.z / sc.z : 0.f, sc.w != 0.f ? is.w / sc.w : 0.f};
    const float mis[4] = {mu.x * is.x, mu.y * is.y, mu.z * is.z, mu.w * is.w};
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    const float inv_keep = g.drop_p > 0.0f ? 1.0f - g.drop_p : 1.0f;
    float a[4] = {0, 0, 0, 0}, bsum[4] = {0, 0, 0, 0};
    float amax = 0.0f;                                                 // max |dz| this thread saw (-> dy_scale_kernel)
    // two pixels per trip: all six loads are issued before the first use
    auto fetch = [&](unsigned pix, float (&av)[4], float (&gq)[4]) {
        unsigned b, h, wo;
        pix_bhw(pix, g, b, h, wo);
        const long off = (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC;
        if (act) {
            load_dA(act + off, g.oC, av);
        } else {
            const float4 t4 = load_planes4(act_hi, act_lo, (long)pix, c4, C4);
            av[0] = t4.x; av[1] = t4.y; av[2] = t4.z; av[3] = t4.w;
        }
        load_dA(dA + off, g.oC, gq);
    };
    auto accumulate = [&](const float (&av)[4], const float (&gq)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (av[q] > 0.0f) {
                const float dz = gq[q] * keep_scale;
                const float xh = fmaf(av[q] * inv_keep - shv[q], rs[q], -mis[q]);
                a[q] += dz;
                bsum[q] = fmaf(dz, xh, bsum[q]);
                amax = fmaxf(amax, fabsf(dz));
            }
        }
    };
    const unsigned stride = gridDim.x * rows;
    unsigned pix = blockIdx.x * rows + prow;
    for (; pix + stride < n_pix; pix += 2 * stride) {
        float av0[4], gq0[4], av1[4], gq1[4];
        fetch(pix, av0, gq0);
        fetch(pix + stride, av1, gq1);
        accumulate(av0, gq0);
        accumulate(av1, gq1);
    }
    if (pix < n_pix) {
        float av0[4], gq0[4];
        fetch(pix, av0, gq0);
        accumulate(av0, gq0);
    }
    s1[threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
    s2[threadIdx.x] = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
    __syncthreads();
    if (prow == 0) {
        float4 ta = make_float4(0, 0, 0, 0), tb = make_float4(0, 0, 0, 0);
        for (int r = 0; r < rows; ++r) {
            const float4 u = s1[r * C4 + c4], w = s2[r * C4 + c4];
            ta.x += u.x; ta.y += u.y; ta.z += u.z; ta.w += u.w;
            tb.x += w.x; tb.y += w.y; tb.z += w.z; tb.w += w.w;
        }
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 0) * g.C + c) = ta;
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 1) * g.C + c) = tb;
    }
    if (amax_part) {                                                   // max is order-independent: still deterministic
        __shared__ float smax[8];
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = amax;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = smax[0];
            for (int w = 1; w < 8; ++w) m = fmaxf(m, smax[w]);
            amax_part[blockIdx.x] = m;
        }
    }
}

// Power-of-two scale of the fp16 gradient plane of one conv block (crnn_block.cuh: store_dy4):
//   |dy| = |sc| * |dz - mean(dz) - xhat * mean(dz*xhat)|  <=  max_c |sc_c| * (max|dz| + |m1_c| + 16 |m2_c|)
// (|xhat| < 16 for any realistic batch; the conversion saturates if that is ever exceeded), placed at 2^13.
// out[0] = scale, out[1] = 1 / scale.
__global__ void __launch_bounds__(128)
dy_scale_kernel(const float* __restrict__ amax_part, int nblk, const float* __restrict__ stat,
                const float* __restrict__ bnsum, int C, float* __restrict__ out) {
    pdl_wait();
    __shared__ float sm[4], sb[4];
    float m = 0.0f, b = 0.0f;
    for (int i = threadIdx.x; i < nblk; i += 128) m = fmaxf(m, amax_part[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    m = fmaxf(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]));
    for (int c = threadIdx.x; c < C; c += 128)
        b = fmaxf(b, fabsf(stat[2 * C + c]) * (m + fabsf(bnsum[c]) + 16.0f * fabsf(bnsum[C + c])));
    b = warp_max(b);
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        b = fmaxf(fmaxf(sb[0], sb[1]), fmaxf(sb[2], sb[3]));
        int e = 0;
        float sc = 1.0f;
        if (b > 0.0f && isfinite(b)) {
            frexpf(b, &e);                                             // b = f * 2^e, f in [0.5, 1)
            sc = exp2f((float)max(-100, min(100, 13 - e)));
        }
        out[0] = sc;
        out[1] = 1.0f / sc;
    }
}

// pass 1: per-channel sum(dz) and sum(dz * xhat); part layout [nblk][2][C]
__global__ void __launch_bounds__(256)
bn_pool_bwd_sums_kernel(const float* __restrict__ y, const float* __restrict__ stat, const float* __restrict__ dA,
                        long n_pix, PoolGeom g, float* __restrict__ part) {
    pdl_wait();
    __shared__ float4 s1[256], s2[256];
    const int C4 = g.C >> 2, rows = 256 / C4;
    const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4;
    const int c = c4 * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
    float a[4] = {0, 0, 0, 0}, bsum[4] = {0, 0, 0, 0};
    for (long pix = (long)blockIdx.x * rows + prow; pix < n_pix; pix += (long)gridDim.x * rows) {
        long t = pix;
        const int wo = (int)(t % g.Wo); t /= g.Wo;
        const int h = (int)(t % g.H);
        const long b = t / g.H;
        const float* src = y + (((b * g.H + h) * g.W) + (long)wo * g.p) * g.C + c;
        const float* da = dA + b * g.oB + h * g.oH + wo * g.oW + c * g.oC;
        const WindowGrad wg = window_grad(src, da, sc, sh, g, pix * C4 + c4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xh = (__ldg(src + (long)wg.arg[q] * g.C + q) - muv[q]) * isv[q];
            a[q] += wg.dz[q];
            bsum[q] = fmaf(wg.dz[q], xh, bsum[q]);
        }
    }
    s1[threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
    s2[threadIdx.x] = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
    __syncthreads();
    if (prow == 0) {
        float4 ta = make_float4(0, 0, 0, 0), tb = make_float4(0, 0, 0, 0);
        for (int r = 0; r < rows; ++r) {
            const float4 u = s1[r * C4 + c4], v = s2[r * C4 + c4];
            ta.x += u.x; ta.y += u.y; ta.z += u.z; ta.w += u.w;
            tb.x += v.x; tb.y += v.y; tb.z += v.z; tb.w += v.w;
        }
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 0) * g.C + c) = ta;
        *reinterpret_cast<float4*>(part + ((long)blockIdx.x * 2 + 1) * g.C + c) = tb;
    }
}

// reduce pass-1 partials: d(beta) = sum dz, d(gamma) = sum dz*xhat; bnsum = {mean dz, mean dz*xhat}
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ part, int nblk, int C, long n,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ bnsum) {
    pdl_wait();
    const int c = blockIdx.x;                                        // one block per channel
    double s, sx;
    reduce_pair(part, nblk, C, c, s, sx);
    if (threadIdx.x != 0) return;
    dbeta[c] = (float)s;
    dgamma[c] = (float)sx;
    bnsum[c] = (float)(s / (double)n);
    bnsum[C + c] = (float)(sx / (double)n);
}

// pass 2: dy = scale * (dz - mean(dz) - xhat * mean(dz*xhat)) for EVERY conv output element
__global__ void __launch_bounds__(256)
bn_pool_bwd_dy_kernel(const float* __restrict__ y, const float* __restrict__ stat, const float* __restrict__ dA,
                      const float* __restrict__ bnsum, long n_vec, PoolGeom g, float* __restrict__ dy,
                      __nv_bfloat16* __restrict__ dy_hi, const float* __restrict__ dy_scale) {
    pdl_wait();
    const int C4 = g.C >> 2;
    const float dscale = dy_scale ? __ldg(dy_scale) : 1.0f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long t = i / C4;
        const int wo = (int)(t % g.Wo); t /= g.Wo;
        const int h = (int)(t % g.H);
        const long b = t / g.H;
        const int c = c4 * 4;
        const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
        const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
        const float4 mu = *reinterpret_cast<const float4*>(stat + c);
        const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
        const float4 k1 = *reinterpret_cast<const float4*>(bnsum + c);
        const float4 k2 = *reinterpret_cast<const float4*>(bnsum + g.C + c);
        const long row = ((b * g.H + h) * g.W) + (long)wo * g.p;
        const float* src = y + row * g.C + c;
        const float* da = dA + b * g.oB + h * g.oH + wo * g.oW + c * g.oC;
        const WindowGrad wg = window_grad(src, da, sc, sh, g, i);
        // the last window of a row also owns the columns the floor-mode pooling drops
        const int span = (wo == g.Wo - 1) ? g.W - wo * g.p : g.p;
        for (int j = 0; j < span; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long)j * g.C));
            float4 o;
            o.x = sc.x * ((wg.arg[0] == j && j < g.p ? wg.dz[0] : 0.0f) - k1.x - (v.x - mu.x) * is.x * k2.x);
            o.y = sc.y * ((wg.arg[1] == j && j < g.p ? wg.dz[1] : 0.0f) - k1.y - (v.y - mu.y) * is.y * k2.y);
            o.z = sc.z * ((wg.arg[2] == j && j < g.p ? wg.dz[2] : 0.0f) - k1.z - (v.z - mu.z) * is.z * k2.z);
            o.w = sc.w * ((wg.arg[3] == j && j < g.p ? wg.dz[3] : 0.0f) - k1.w - (v.w - mu.w) * is.w * k2.w);
            if (dy) *reinterpret_cast<float4*>(dy + (row + j) * g.C + c) = o;
            if (dy_hi) store_dy4(dy_hi, ((row + j) * g.C + c) >> 2, o, dscale);
        }
    }
}


// Same pass with a compile-time pool width that divides W (no floor-dropped columns): the P conv outputs of a
// window are loaded once, the winner is found in registers, and the pixel index uses 32-bit arithmetic.
//   dy = scale*dz[winner] - A - Bc*y   with   Bc = scale*invstd*k2,  A = scale*k1 - mean*Bc
template <int P>
__global__ void __launch_bounds__(256)
bn_pool_bwd_dy_t_kernel(const float* __restrict__ y, const float* __restrict__ stat, const float* __restrict__ dA,
                        const float* __restrict__ bnsum, unsigned n_pix, PoolGeom g, float* __restrict__ dy,
                        __nv_bfloat16* __restrict__ dy_hi, const float* __restrict__ dy_scale) {
    pdl_wait();
    const int C4 = g.C >> 2, rows = 256 / C4;
    const float dscale = dy_scale ? __ldg(dy_scale) : 1.0f;
    const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4, c = c4 * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    const float4 k1 = *reinterpret_cast<const float4*>(bnsum + c);
    const float4 k2 = *reinterpret_cast<const float4*>(bnsum + g.C + c);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    const float Bc[4] = {sc.x * is.x * k2.x, sc.y * is.y * k2.y, sc.z * is.z * k2.z, sc.w * is.w * k2.w};
    const float Ac[4] = {sc.x * k1.x - mu.x * Bc[0], sc.y * k1.y - mu.y * Bc[1], sc.z * k1.z - mu.z * Bc[2],
                         sc.w * k1.w - mu.w * Bc[3]};
    for (unsigned pix = blockIdx.x * rows + prow; pix < n_pix; pix += gridDim.x * rows) {
        unsigned b, h, wo;
        pix_bhw(pix, g, b, h, wo);
        const long row = ((long)b * g.H + h) * g.W + (long)wo * P;
        float4 v[P];
        float gq[4], dz[4], yarg[4];
        int arg[4];
        load_window<P>(y + row * g.C + c, g.C, v);
        load_dA(dA + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC, g.oC, gq);
        eval_window<P>(v, scv, shv, gq, g, (long)pix * C4 + c4, dz, arg, yarg);
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fmaf(-Bc[q], vv[q], (arg[q] == j ? scv[q] * dz[q] : 0.0f) - Ac[q]);
            const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
            if (dy) *reinterpret_cast<float4*>(dy + (row + j) * g.C + c) = o4;
            if (dy_hi) store_dy4(dy_hi, ((row + j) * g.C + c) >> 2, o4, dscale);
        }
    }
}


// ----------------------------------------------------------------------------- first conv block, direct
// conv block 0 has K = 9*Cin = 9 or 18: a degenerate GEMM whose cost is the 4 B x B*H*W*C activation it
// produces, not its FLOPs.  Two direct fp32 kernels keep that tensor's HBM traffic at the minimum:
//   forward : conv + bias -> y0 (one write) with the BatchNorm statistics accumulated on the fly
//   backward: BN/ReLU/pool backward recomputed from y0 and consumed IN REGISTERS by the weight-gradient
//             accumulation -- dy0 is never written (conv 0 needs no data gradient).
// Thread mapping: lane <-> 4 consecutive output channels (weights / dW live in registers for the whole
// kernel), warp <-> one image row, block <-> 8 rows; the input rows (with zero halo) sit in shared memory
// and are read as warp-wide broadcasts; every y0 access is a 512 B coalesced row of 128 channels.

template <int CIN>
__device__ __forceinline__ void load_x_rows(const float* __restrict__ x, float* xs, int b, int h0, int H, int W) {
    const int Wp = W + 2;
    for (int idx = threadIdx.x; idx < CIN * (kC0Rows + 2) * Wp; idx += blockDim.x) {
        const int cc = idx % Wp, rr = (idx / Wp) % (kC0Rows + 2), ci = idx / (Wp * (kC0Rows + 2));
        const int hh = h0 - 1 + rr, ww = cc - 1;
        xs[idx] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (((long)b * CIN + ci) * H + hh) * W + ww) : 0.0f;
    }
}

// persistent: a block walks row groups (8 image rows) with stride gridDim.x, BatchNorm partial sums stay in
// registers across groups (one partial per block), next group's input rows are fetched while computing
template <int CIN>
__global__ void __launch_bounds__(256)
conv0_fwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ y, int H, int W, int C, int groups_per_img, int n_groups,
                       float* __restrict__ part) {
    pdl_wait();
    extern __shared__ float xs_all[];                     // 2 x [CIN][10][W+2]
    __shared__ float red[kC0Rows][2][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.y * 128 + lane * 4, Wp = W + 2, xsz = CIN * (kC0Rows + 2) * Wp;
    // weights as channel pairs: one packed fma.f32x2 updates two output channels (halves the FMA issue slots)
    float2 wr[2][CIN * 9], bs[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        bs[q] = make_float2(__ldg(bias + c + 2 * q), __ldg(bias + c + 2 * q + 1));
#pragma unroll
        for (int k = 0; k < CIN * 9; ++k)
            wr[q][k] = make_float2(__ldg(w + (long)(c + 2 * q) * CIN * 9 + k), __ldg(w + (long)(c + 2 * q + 1) * CIN * 9 + k));
    }
    float2 s1[2] = {make_float2(0, 0), make_float2(0, 0)}, s2[2] = {make_float2(0, 0), make_float2(0, 0)};
    int grp = blockIdx.x, buf = 0;
    if (grp < n_groups) load_x_rows<CIN>(x, xs_all, grp / groups_per_img, (grp % groups_per_img) * kC0Rows, H, W);
    __syncthreads();
    for (; grp < n_groups; grp += gridDim.x, buf ^= 1) {
        const float* xs = xs_all + buf * xsz;
        const int b = grp / groups_per_img, h0 = (grp % groups_per_img) * kC0Rows;
        const int nxt = grp + gridDim.x;
        if (nxt < n_groups)
            load_x_rows<CIN>(x, xs_all + (buf ^ 1) * xsz, nxt / groups_per_img, (nxt % groups_per_img) * kC0Rows, H, W);
        const int h = h0 + warp;
        if (h < H) {
            float* yrow = y + ((long)b * H + h) * W * C + c;
            // two neighbouring pixels per trip: four independent accumulator chains, and the 4 input columns they
            // share are read once (12 instead of 18 shared-memory broadcasts per pixel)
            int ww = 0;
            for (; ww + 1 < W; ww += 2) {
                float2 acc[2][2] = {{bs[0], bs[1]}, {bs[0], bs[1]}};
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float* xr = xs + (ci * (kC0Rows + 2) + warp + r) * Wp + ww;
                        float xc[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) xc[t] = xr[t];
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const float2 x0 = make_float2(xc[t], xc[t]), x1 = make_float2(xc[t + 1], xc[t + 1]);
                            acc[0][0] = __ffma2_rn(wr[0][ci * 9 + r * 3 + t], x0, acc[0][0]);
                            acc[0][1] = __ffma2_rn(wr[1][ci * 9 + r * 3 + t], x0, acc[0][1]);
                            acc[1][0] = __ffma2_rn(wr[0][ci * 9 + r * 3 + t], x1, acc[1][0]);
                            acc[1][1] = __ffma2_rn(wr[1][ci * 9 + r * 3 + t], x1, acc[1][1]);
                        }
                    }
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    *reinterpret_cast<float4*>(yrow + (long)(ww + px) * C) =
                        make_float4(acc[px][0].x, acc[px][0].y, acc[px][1].x, acc[px][1].y);
#pragma unroll
                    for (int q = 0; q < 2; ++q) { s1[q] = __fadd2_rn(s1[q], acc[px][q]); s2[q] = __ffma2_rn(acc[px][q], acc[px][q], s2[q]); }
                }
            }
            for (; ww < W; ++ww) {
                float2 acc[2] = {bs[0], bs[1]};
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const float xv = xs[(ci * (kC0Rows + 2) + warp + r) * Wp + ww + t];
                            const float2 x2 = make_float2(xv, xv);
                            acc[0] = __ffma2_rn(wr[0][ci * 9 + r * 3 + t], x2, acc[0]);
                            acc[1] = __ffma2_rn(wr[1][ci * 9 + r * 3 + t], x2, acc[1]);
                        }
                *reinterpret_cast<float4*>(yrow + (long)ww * C) = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
#pragma unroll
                for (int q = 0; q < 2; ++q) { s1[q] = __fadd2_rn(s1[q], acc[q]); s2[q] = __ffma2_rn(acc[q], acc[q], s2[q]); }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        red[warp][0][lane * 4 + 2 * q] = s1[q].x; red[warp][0][lane * 4 + 2 * q + 1] = s1[q].y;
        red[warp][1][lane * 4 + 2 * q] = s2[q].x; red[warp][1][lane * 4 + 2 * q + 1] = s2[q].y;
    }
    __syncthreads();
    {
        const int which = threadIdx.x >> 7, ch = threadIdx.x & 127;
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < kC0Rows; ++r) t += red[r][which][ch];
        part[((long)blockIdx.x * 2 + which) * C + blockIdx.y * 128 + ch] = t;
    }
}

// part layout: [nblk][CIN*9 + 1][C]   (slot CIN*9 = bias gradient)
template <int CIN>
__global__ void __launch_bounds__(256)
conv0_bwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ stat,
                       const float* __restrict__ dA, const float* __restrict__ bnsum, PoolGeom g,
                       float* __restrict__ part) {
    pdl_wait();
    extern __shared__ float xs[];
    __shared__ float red[kC0Rows][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, h0 = blockIdx.x * kC0Rows, c4 = blockIdx.z * 32 + lane, c = c4 * 4, Wp = g.W + 2;
    const int C4 = g.C >> 2;
    load_x_rows<CIN>(x, xs, b, h0, g.H, g.W);
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    const float4 k1 = *reinterpret_cast<const float4*>(bnsum + c);
    const float4 k2 = *reinterpret_cast<const float4*>(bnsum + g.C + c);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, muv[4] = {mu.x, mu.y, mu.z, mu.w};
    const float isv[4] = {is.x, is.y, is.z, is.w}, k1v[4] = {k1.x, k1.y, k1.z, k1.w}, k2v[4] = {k2.x, k2.y, k2.z, k2.w};
    float dw[4][CIN * 9 + 1];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k <= CIN * 9; ++k) dw[q][k] = 0.0f;
    __syncthreads();
    const int h = h0 + warp;
    if (h < g.H) {
        for (int wo = 0; wo < g.Wo; ++wo) {
            const long row = (((long)b * g.H + h) * g.W) + (long)wo * g.p;
            const float* src = y + row * g.C + c;
            const float* da = dA + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC;
            const long i = (((long)b * g.H + h) * g.Wo + wo) * C4 + c4;
            const WindowGrad wg = window_grad(src, da, sc, sh, g, i);
            const int span = (wo == g.Wo - 1) ? g.W - wo * g.p : g.p;
            for (int j = 0; j < span; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long)j * g.C));
                const float vv[4] = {v.x, v.y, v.z, v.w};
                float dyv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    dyv[q] = scv[q] * ((wg.arg[q] == j && j < g.p ? wg.dz[q] : 0.0f) - k1v[q] - (vv[q] - muv[q]) * isv[q] * k2v[q]);
                const int ww = wo * g.p + j;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const float xv = xs[(ci * (kC0Rows + 2) + warp + r) * Wp + ww + t];
#pragma unroll
                            for (int q = 0; q < 4; ++q) dw[q][ci * 9 + r * 3 + t] = fmaf(dyv[q], xv, dw[q][ci * 9 + r * 3 + t]);
                        }
#pragma unroll
                for (int q = 0; q < 4; ++q) dw[q][CIN * 9] += dyv[q];
            }
        }
    }
    const long blk = (long)blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
    for (int k = 0; k <= CIN * 9; ++k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) red[warp][lane * 4 + q] = dw[q][k];
        __syncthreads();
        if (threadIdx.x < 128) {
            float t = 0.0f;
#pragma unroll
            for (int r = 0; r < kC0Rows; ++r) t += red[r][threadIdx.x];
            part[(blk * (CIN * 9 + 1) + k) * g.C + blockIdx.z * 128 + threadIdx.x] = t;
        }
        __syncthreads();
    }
}


// Same as conv0_bwd_fused_kernel with the pool width P known at compile time and W == Wo*P, PERSISTENT: a block
// walks row groups (8 image rows each) with a stride of gridDim.x, keeps its dW / db accumulators in registers
// across all of them and reduces once at the end (gridDim.x partials instead of one per row group).  The input
// rows of the next group are fetched into the second shared buffer while the current group is consumed, and
// the P conv outputs of the next pooling window are in flight while the current window is processed.
template <int CIN, int P>
__global__ void __launch_bounds__(256)
conv0_bwd_fused_t_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ stat,
                         const float* __restrict__ dA, const float* __restrict__ bnsum, PoolGeom g, int groups_per_img,
                         int n_groups, float* __restrict__ part) {
    pdl_wait();
    extern __shared__ float xs_all[];                 // 2 x [CIN][10][W+2]
    __shared__ float red[kC0Rows][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = blockIdx.y * 32 + lane, c = c4 * 4, Wp = g.W + 2;
    const int C4 = g.C >> 2, xsz = CIN * (kC0Rows + 2) * Wp;
    const float4 sc = *reinterpret_cast<const float4*>(stat + 2 * g.C + c);
    const float4 sh = *reinterpret_cast<const float4*>(stat + 3 * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(stat + c);
    const float4 is = *reinterpret_cast<const float4*>(stat + g.C + c);
    const float4 k1 = *reinterpret_cast<const float4*>(bnsum + c);
    const float4 k2 = *reinterpret_cast<const float4*>(bnsum + g.C + c);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    // dy = scale*dz - A - B*y   with  A = scale*(k1 - mu*is*k2),  B = scale*is*k2
    const float Bc[4] = {sc.x * is.x * k2.x, sc.y * is.y * k2.y, sc.z * is.z * k2.z, sc.w * is.w * k2.w};
    const float Ac[4] = {sc.x * k1.x - mu.x * Bc[0], sc.y * k1.y - mu.y * Bc[1], sc.z * k1.z - mu.z * Bc[2],
                         sc.w * k1.w - mu.w * Bc[3]};
    float2 dw[2][CIN * 9 + 1];                        // channel pairs: packed fma.f32x2
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k <= CIN * 9; ++k) dw[q][k] = make_float2(0.0f, 0.0f);

    int grp = blockIdx.x, buf = 0;
    if (grp < n_groups) load_x_rows<CIN>(x, xs_all, grp / groups_per_img, (grp % groups_per_img) * kC0Rows, g.H, g.W);
    __syncthreads();
    for (; grp < n_groups; grp += gridDim.x, buf ^= 1) {
        const float* xs = xs_all + buf * xsz;
        const int b = grp / groups_per_img, h0 = (grp % groups_per_img) * kC0Rows;
        const int nxt = grp + gridDim.x;
        if (nxt < n_groups)                           // next group's input rows -> the other buffer
            load_x_rows<CIN>(x, xs_all + (buf ^ 1) * xsz, nxt / groups_per_img, (nxt % groups_per_img) * kC0Rows, g.H, g.W);
        const int h = h0 + warp;
        if (h < g.H) {
            const float* yrow = y + (((long)b * g.H + h) * g.W) * g.C + c;
            const float* darow = dA + (long)b * g.oB + (long)h * g.oH + (long)c * g.oC;
            float4 v[P], vn[P];
            float gq[4], gn[4];
            load_window<P>(yrow, g.C, v);
            load_dA(darow, g.oC, gq);
            for (int wo = 0; wo < g.Wo; ++wo) {
                if (wo + 1 < g.Wo) {
                    load_window<P>(yrow + (long)(wo + 1) * P * g.C, g.C, vn);
                    load_dA(darow + (long)(wo + 1) * g.oW, g.oC, gn);
                }
                float dz[4], yarg[4];
                int arg[4];
                eval_window<P>(v, scv, shv, gq, g, (((long)b * g.H + h) * g.Wo + wo) * C4 + c4, dz, arg, yarg);
                // dy of the window's P pixels first, then one pass per (input channel, kernel row) in which the P + 2
                // input columns the window touches are read ONCE (42 instead of 90 shared-memory broadcasts for P = 5);
                // every dw element still receives its P contributions in ascending pixel order
                float2 d01[P], d23[P];
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                    float dyv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) dyv[q] = fmaf(-Bc[q], vv[q], (arg[q] == j ? scv[q] * dz[q] : 0.0f) - Ac[q]);
                    d01[j] = make_float2(dyv[0], dyv[1]);
                    d23[j] = make_float2(dyv[2], dyv[3]);
                    dw[0][CIN * 9] = __fadd2_rn(dw[0][CIN * 9], d01[j]);
                    dw[1][CIN * 9] = __fadd2_rn(dw[1][CIN * 9], d23[j]);
                }
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float* xr = xs + (ci * (kC0Rows + 2) + warp + r) * Wp + wo * P;
                        float xc[P + 2];
#pragma unroll
                        for (int t = 0; t < P + 2; ++t) xc[t] = xr[t];
#pragma unroll
                        for (int t = 0; t < 3; ++t)
#pragma unroll
                            for (int j = 0; j < P; ++j) {
                                const float2 x2 = make_float2(xc[j + t], xc[j + t]);
                                dw[0][ci * 9 + r * 3 + t] = __ffma2_rn(d01[j], x2, dw[0][ci * 9 + r * 3 + t]);
                                dw[1][ci * 9 + r * 3 + t] = __ffma2_rn(d23[j], x2, dw[1][ci * 9 + r * 3 + t]);
                            }
                    }
#pragma unroll
                for (int j = 0; j < P; ++j) v[j] = vn[j];
#pragma unroll
                for (int q = 0; q < 4; ++q) gq[q] = gn[q];
            }
        }
        __syncthreads();                              // everyone is done with xs[buf]; xs[buf^1] is complete
    }
#pragma unroll
    for (int k = 0; k <= CIN * 9; ++k) {
        red[warp][lane * 4 + 0] = dw[0][k].x; red[warp][lane * 4 + 1] = dw[0][k].y;
        red[warp][lane * 4 + 2] = dw[1][k].x; red[warp][lane * 4 + 3] = dw[1][k].y;
        __syncthreads();
        if (threadIdx.x < 128) {
            float t = 0.0f;
#pragma unroll
            for (int r = 0; r < kC0Rows; ++r) t += red[r][threadIdx.x];
            part[((long)blockIdx.x * (CIN * 9 + 1) + k) * g.C + blockIdx.y * 128 + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// dW[c][j] = sum_blk part[blk][j][c] (j < J), db[c] = sum_blk part[blk][J][c]; one warp per output
__global__ void conv0_bwd_reduce_kernel(const float* __restrict__ part, int nblk, int J, int C,
                                        float* __restrict__ dw, float* __restrict__ db) {
    pdl_wait();
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= (J + 1) * C) return;
    const int j = o / C, c = o - j * C;
    double a = 0.0;
    for (int k = lane; k < nblk; k += 32) a += (double)__ldg(part + ((long)k * (J + 1) + j) * C + c);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) {
        if (j < J) dw[(long)c * J + j] = (float)a;
        else db[c] = (float)a;
    }
}


// ----------------------------------------------------------------------------- small dense layers, backward
// The per-frame dense head is tiny (e.g. 64 -> 16 -> 6): one kernel per layer produces d(input) for its 128
// rows and the block's partial dW / db; a second one sums the partials in a fixed order.
constexpr int kDbRows = 128;
__global__ void __launch_bounds__(256)
dense_bwd_small_kernel(const float* __restrict__ dout, const float* __restrict__ in, const float* __restrict__ W,
                       const float* __restrict__ relu_act /* activation of the producing layer or null */,
                       int rows, int N, int D, float* __restrict__ din, float* __restrict__ part) {
    pdl_wait();
    extern __shared__ float sm[];
    float* s_do = sm;                       // [128][N]
    float* s_in = s_do + kDbRows * N;       // [128][D+1]
    float* s_w = s_in + kDbRows * (D + 1);  // [N][D]
    const int r0 = blockIdx.x * kDbRows, nr = min(kDbRows, rows - r0), Dp = D + 1;
    for (int i = threadIdx.x; i < kDbRows * N; i += 256) s_do[i] = (i / N) < nr ? __ldg(dout + (long)r0 * N + i) : 0.0f;
    for (int i = threadIdx.x; i < kDbRows * D; i += 256) {
        const int r = i / D, k = i - r * D;
        s_in[r * Dp + k] = r < nr ? __ldg(in + (long)r0 * D + i) : 0.0f;
    }
    for (int i = threadIdx.x; i < N * D; i += 256) s_w[i] = __ldg(W + i);
    __syncthreads();
    // d(in)[r][k] = sum_n dout[r][n] W[n][k]
    if (din) {
        for (int i = threadIdx.x; i < nr * D; i += 256) {
            const int r = i / D, k = i - r * D;
            float acc = 0.0f;
            for (int n = 0; n < N; ++n) acc = fmaf(s_do[r * N + n], s_w[n * D + k], acc);
            if (relu_act && !(__ldg(relu_act + (long)(r0 + r) * D + k) > 0.0f)) acc = 0.0f;
            din[(long)(r0 + r) * D + k] = acc;
        }
    }
    // partial dW[n][k] = sum_r dout[r][n] in[r][k];  partial db[n] = sum_r dout[r][n]
    float* pb = part + (long)blockIdx.x * (N * D + N);
    for (int i = threadIdx.x; i < N * D; i += 256) {
        const int n = i / D, k = i - n * D;
        float acc = 0.0f;
        for (int r = 0; r < kDbRows; ++r) acc = fmaf(s_do[r * N + n], s_in[r * Dp + k], acc);
        pb[i] = acc;
    }
    for (int n = threadIdx.x; n < N; n += 256) {
        float acc = 0.0f;
        for (int r = 0; r < kDbRows; ++r) acc += s_do[r * N + n];
        pb[N * D + n] = acc;
    }
}
__global__ void dense_bwd_reduce_kernel(const float* __restrict__ part, int nblk, int ND, int N,
                                        float* __restrict__ dw, float* __restrict__ db) {
    pdl_wait();
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= ND + N) return;
    double a = 0.0;
    for (int k = lane; k < nblk; k += 32) a += (double)__ldg(part + (long)k * (ND + N) + o);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) {
        if (o < ND) dw[o] = (float)a;
        else db[o - ND] = (float)a;
    }
}
inline size_t dense_small_smem(int N, int D) { return ((size_t)kDbRows * N + (size_t)kDbRows * (D + 1) + (size_t)N * D) * 4; }
inline bool dense_small_ok(int N, int D) { return dense_small_smem(N, D) <= 160 * 1024; }

// part_b [B][2][n6] -> dbih[n6], dbhh[n6]  (fixed-order sum over B)
__global__ void reduce_bias_partials_kernel(const float* __restrict__ part_b, float* __restrict__ dbih,
                                            float* __restrict__ dbhh, int n6, int B) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n6) return;
    const float acc = ordered_sum<16, float>(part_b + i, 2L * n6, B);
    if (i < n6) dbih[i] = acc;
    else dbhh[i - n6] = acc;
}
// whh[dir][r][j] = full[dir*3h + r][dir*h + j],  full is [6h][2h]
__global__ void extract_whh_kernel(const float* __restrict__ full, int h, float* __restrict__ whh) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * h * h) return;
    const int j = i % h, r = (i / h) % (3 * h), dir = i / (3 * h * h);
    whh[i] = full[(long)(dir * 3 * h + r) * 2 * h + dir * h + j];
}
inline int reduce_bias_partials(const float* part_b, float* dbih, float* dbhh, int n6, int B, cudaStream_t st) {
    launch_k(reduce_bias_partials_kernel, (2 * n6 + 127) / 128, 128, 0, st, part_b, dbih, dbhh, n6, B);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

inline int ew_blocks(long n) { return (int)std::min<long>((n + 255) / 256, 148L * 16); }

PoolGeom pool_geom(const Plan& P, const sedb200_crnn_desc* d, int i, int training, unsigned long long seed,
                   const unsigned long long* seed_ptr = nullptr) {
    PoolGeom g;
    g.seed_ptr = seed_ptr; g.block = i;
    g.H = P.H; g.W = P.win[i]; g.Wo = P.wout[i]; g.C = P.C; g.p = P.pool[i];
    const bool last = (i == P.n_conv - 1);
    if (!last) {
        g.oC = 1; g.oW = P.C; g.oH = (long)g.Wo * P.C; g.oB = (long)P.H * g.Wo * P.C;
    } else if (d->mode == 0) {          // [B][T=wo][c*H + h]
        g.oB = (long)P.T * P.flat; g.oW = P.flat; g.oC = P.H; g.oH = 1;
    } else {                            // [B][T=h][c*Wo + wo]
        g.oB = (long)P.T * P.flat; g.oH = P.flat; g.oC = g.Wo; g.oW = 1;
    }
    const bool drop = training && d->dropout > 0.0f && (d->dropout_each_block || last);
    g.drop_p = drop ? d->dropout : 0.0f;
    g.seed = block_seed(seed, i);
    auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    g.lgWo = lg(g.Wo); g.lgH = lg(g.H);
    if (g.lgWo < 0 || g.lgH < 0) g.lgWo = g.lgH = -1;
    return g;
}

struct InStrides { long sB, sH, sW, sC; };
InStrides in_strides(const Plan& P, const sedb200_crnn_desc* d, int i) {
    if (i == 0) return {(long)d->in_ch * P.H * P.win[0], (long)P.win[0], 1L, (long)P.H * P.win[0]};   // NCHW
    return {(long)P.H * P.win[i] * P.C, (long)P.win[i] * P.C, (long)P.C, 1L};                          // NHWC
}

inline float* wsf(void* ws, size_t off) { return reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + off); }

// One helper stream + fork / done events per device (created once, never destroyed): the GRU backward pass runs the
// weight-gradient GEMMs of a layer there while the caller's stream continues with the critical path.
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr}, done[4] = {nullptr, nullptr, nullptr, nullptr};
};                                                           // [0..1]: GRU layers by parity, [2..3]: conv blocks by parity
std::mutex g_side_mu;
std::map<int, SideStream> g_side;
int side_stream(SideStream** out) {
    int dev = 0;
    SED_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    auto it = g_side.find(dev);
    if (it == g_side.end()) {
        SideStream s;
        SED_CUDA_OK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        pdl_exclude_stream(s.st);
        for (int i = 0; i < 4; ++i) {
            SED_CUDA_OK(cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming));
            SED_CUDA_OK(cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming));
        }
        it = g_side.emplace(dev, s).first;
    }
    *out = &it->second;
    return SEDB200_OK;
}


// keep-mask of one block's dropout as the reference would hold it: NCHW bytes [B][C][H][Wo] (1 = kept), from the same
// counter-based generator the fused kernels evaluate (element group i = ((b*H + h)*Wo + wo)*C/4 + c/4)
__global__ void __launch_bounds__(256)
dropout_mask_kernel(unsigned char* __restrict__ mask, long n_vec, PoolGeom g) {
    pdl_wait();
    const int C4 = g.C >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long t = i / C4;
        const int wo = (int)(t % g.Wo); t /= g.Wo;
        const int h = (int)(t % g.H);
        const long b = t / g.H;
        Keep4 kp;
#pragma unroll
        for (int q = 0; q < 4; ++q) kp.k[q] = true;
        if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            mask[((b * g.C + (c4 * 4 + q)) * g.H + h) * g.Wo + wo] = kp.k[q] ? 1 : 0;
    }
}
}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

// programmatic dependent launch inside a captured step only for models without tensor-core conv blocks: those steps are
// chains of ~10 us kernels (launch-bound); the large models keep the concurrency of their forked branches instead
static bool plan_launch_bound(const Plan& P) {
    for (int i = 0; i < P.n_conv; ++i) if (P.conv_tc_all[i]) return false;
    return true;
}

static int crnn_forward_impl(const sedb200_crnn_desc* d, const float* params, float* bn_state, const float* x,
                             int batch, int training, unsigned long long seed, const unsigned long long* seed_ptr,
                             void* ws, size_t ws_bytes, float* logits, void* stream) {
    Plan P;
    int rc = make_plan(d, batch, &P);
    if (rc) return rc;
    PdlCaptureScope pdl_scope(plan_launch_bound(P));
    SED_REQUIRE(batch >= 1, SEDB200_EINVAL, "crnn_forward: batch %d", batch);
    SED_REQUIRE(params && bn_state && x && ws, SEDB200_EINVAL, "crnn_forward: null buffer");
    SED_REQUIRE(ws_bytes >= P.ws_bytes, SEDB200_EWORKSPACE, "crnn_forward: workspace %zu < %zu bytes", ws_bytes, P.ws_bytes);
    rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const long B = batch;

    // ---- operand planes of the weights (conv blocks on tcgen05: forward and flipped data-gradient layout; W_ih of the
    //      tensor-core GRU layers), built on the side stream while block 0 runs; the backward pass reuses them.  The fork
    //      point is the start of the call (the parameters are final there); the side-stream launches themselves are
    //      enqueued AFTER block 0's own kernels, so that a caller who synchronises every step sees the GPU busy sooner.
    SideStream* wside = nullptr;
    bool wp_pending = false, wp_forked = false;
    if (P.weight_planes_ahead) {
        rc = side_stream(&wside);
        if (rc) return rc;
        SED_CUDA_OK(cudaEventRecord(wside->fork[0], st));
    }
    auto enqueue_weight_planes = [&]() -> int {
        if (!P.weight_planes_ahead || wp_forked) return SEDB200_OK;
        wp_forked = true;
        cudaStream_t ss = wside->st;
        SED_CUDA_OK(cudaStreamWaitEvent(ss, wside->fork[0], 0));
        for (int i = 1; i < P.n_conv; ++i) {
            if (!P.conv_tc_all[i]) continue;
            for (int dg = 0; dg < 2; ++dg) {
                // forward layout: fp16 hi + the combined e4m3 correction plane and its scale; data-gradient layout: fp16 hi
                const int rc2 = conv_tc_weight_planes(params + P.conv_w[i], P.cin[i], P.C, dg,
                                                      reinterpret_cast<char*>(ws) + P.wpl[i][dg], ss, kPlaneF16,
                                                      dg == 0 ? wsf(ws, P.dys) + 1040 + 2 * i : nullptr);
                if (rc2) return rc2;
            }
        }
        for (int l = 0; l < P.n_gru; ++l) {
            if (!P.gru_tc[l]) continue;
            const size_t wpb = ((size_t)6 * P.gh[l] * P.gin[l] * 2 + 1023) & ~(size_t)1023;
            char* wp = reinterpret_cast<char*>(ws) + P.wihp[l];
            const int rc2 = split_planes(params + P.wih[l], wp, wp + wpb, 6L * P.gh[l] * P.gin[l], ss);
            if (rc2) return rc2;
        }
        SED_CUDA_OK(cudaEventRecord(wside->done[0], ss));
        wp_pending = true;
        return SEDB200_OK;
    };
    struct WpJoin {             // rejoin on every way out, and before the first consumer
        SideStream*& s; bool& pending; cudaStream_t st;
        void now() { if (pending) { cudaStreamWaitEvent(st, s->done[0], 0); pending = false; } }
        ~WpJoin() { now(); }
    } wp_join{wside, wp_pending, st};

    // ---- conv blocks
    for (int i = 0; i < P.n_conv; ++i) {
        if (i > 0) {
            rc = enqueue_weight_planes();
            if (rc) return rc;
            wp_join.now();
        }
        const float* in = i == 0 ? x : wsf(ws, P.act[i - 1]);
        const InStrides s = in_strides(P, d, i);
        const int M = (int)(B * P.H * P.win[i]), K = 9 * P.cin[i];
        float* y = wsf(ws, P.y[i]);
        float* stat = wsf(ws, P.stat[i]);
        float* running = bn_state + 2L * i * P.C;
        const bool direct0 = (i == 0) && conv0_direct_ok(P.cin[0], P.C);
        int nblk = 0;
        if (i == 0 && conv0_lean_ok(P.cin[0], P.C, P.pool[0], P.n_conv)) {
            // lean block 0 (conv0_lean.cu): statistics from the patch moments of the input, then one fused kernel; y0 is
            // never stored
            {
                SED_PROF("conv0.stats", st);
                if (training) {
                    rc = conv0_lean_stats(x, P.cin[0], P.C, P.H, P.win[0], batch, params + P.conv_w[0], params + P.conv_b[0],
                                          params + P.bn_w[0], params + P.bn_b[0], d->bn_eps, d->bn_momentum, running, stat,
                                          reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + P.gram), wsf(ws, P.part), st);
                    if (rc) return rc;
                } else {
                    launch_k(bn_finalize_eval_kernel, (P.C + 127) / 128, 128, 0, st, P.C, params + P.bn_w[0], params + P.bn_b[0],
                                                                               d->bn_eps, running, stat);
                    SED_POST_LAUNCH();
                }
            }
            SED_PROF("conv0.fwd_fused", st);
            const PoolGeom g = pool_geom(P, d, 0, training, seed, seed_ptr);
            const bool to_planes = (P.n_conv > 1) && P.conv_tc_all[1];
            __nv_bfloat16* ph = to_planes ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + P.actp[0]) : nullptr;
            __nv_bfloat16* pl = to_planes ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + P.actp[0] + P.act_plane_bytes[0]) : nullptr;
            float* of = to_planes ? nullptr : wsf(ws, P.act[0]);
            unsigned* argw = training ? reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + P.arg0) : nullptr;
            rc = conv0_lean_forward(x, P.cin[0], batch, params + P.conv_w[0], params + P.conv_b[0], stat, g, d->tensor_cores,
                                    of, ph, pl, argw, st);
            if (rc) return rc;
            continue;
        }
        { char _nm[40]; snprintf(_nm, sizeof _nm, "conv%d.fwd", i); SED_PROF(_nm, st);
        if (direct0) {
            const int gpi = (P.H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch;
            const dim3 grid(std::min(n_groups, 2 * sm_count()), P.C / 128);
            const size_t sm = 2 * (size_t)P.cin[0] * (kC0Rows + 2) * (P.win[0] + 2) * 4;
            if (P.cin[0] == 1)
                launch_k(conv0_fwd_stats_kernel<1>, grid, 256, sm, st, x, params + P.conv_w[0], params + P.conv_b[0], y, P.H, P.win[0], P.C, gpi, n_groups, wsf(ws, P.part));
            else
                launch_k(conv0_fwd_stats_kernel<2>, grid, 256, sm, st, x, params + P.conv_w[0], params + P.conv_b[0], y, P.H, P.win[0], P.C, gpi, n_groups, wsf(ws, P.part));
            SED_POST_LAUNCH();
            nblk = (int)grid.x;
        } else if (P.conv_tc_all[i]) {
            // plane-native: input planes were written by the previous block's pool kernel; BatchNorm partial sums
            // come out of the conv epilogue
            const char* ap = reinterpret_cast<const char*>(ws) + P.actp[i - 1];
            rc = conv_tc_planes_w(ap, ap + P.act_plane_bytes[i - 1], reinterpret_cast<const char*>(ws) + P.wpl[i][0],
                                  params + P.conv_b[i], y, training ? wsf(ws, P.part) : nullptr, batch, P.H, P.win[i],
                                  P.cin[i], P.C, 0, st, 2, kPlaneF16, nullptr, wsf(ws, P.dys) + 1040 + 2 * i + 1);
            if (rc) return rc;
            nblk = conv_tc_stat_tiles(batch, P.H, P.win[i]);
        } else if (i > 0 && d->tensor_cores && conv_tc_supported(P.H, P.win[i], P.cin[i], P.C)) {
            rc = conv_tc_forward(in, params + P.conv_w[i], params + P.conv_b[i], y, batch, P.H, P.win[i], P.cin[i], P.C,
                                 0, wsf(ws, P.tc), P.tc_bytes, st);
            if (rc) return rc;
        } else if (conv_small_supported(P.cin[i], P.C, P.win[i])) {
            rc = conv_small_forward(in, s.sB, s.sH, s.sW, s.sC, P.cin[i], batch, P.H, P.win[i], params + P.conv_w[i],
                                    params + P.conv_b[i], P.C, 0, y, st);
            if (rc) return rc;
        } else {
            rc = gemm_simt(M, P.C, K, 1, ConvFwdA{in, P.H, P.win[i], P.cin[i], s.sB, s.sH, s.sW, s.sC},
                           ConvFwdB{params + P.conv_w[i], P.cin[i]}, EpiStore{y, P.C, params + P.conv_b[i], 0}, st);
            if (rc) return rc;
        }
        }
        { char _nm[40]; snprintf(_nm, sizeof _nm, "bn%d.stats", i); SED_PROF(_nm, st);
        if (training) {
            if (!direct0 && !P.conv_tc_all[i]) {
                rc = colsum_partials(y, M, P.C, wsf(ws, P.part), &nblk, st);
                if (rc) return rc;
            }
            launch_k(bn_finalize_train_kernel, P.C, 128, 0, st,
                wsf(ws, P.part), nblk, P.C, (long)M, params + P.bn_w[i], params + P.bn_b[i], d->bn_eps,
                d->bn_momentum, running, stat);
        } else {
            launch_k(bn_finalize_eval_kernel, (P.C + 127) / 128, 128, 0, st, P.C, params + P.bn_w[i], params + P.bn_b[i],
                                                                       d->bn_eps, running, stat);
        }
        SED_POST_LAUNCH();
        }
        const PoolGeom g = pool_geom(P, d, i, training, seed, seed_ptr);
        const long n_vec = B * P.H * P.wout[i] * (P.C / 4);
{ char _nm[40]; snprintf(_nm, sizeof _nm, "pool%d.fwd", i); SED_PROF(_nm, st);
        {
            // the consumer decides the format: a plane-native conv wants bf16 hi/lo planes only, the GRU wants the
            // fp32 sequence (plus planes of it when its projection runs on tcgen05), everything else fp32
            const bool to_planes = (i + 1 < P.n_conv) && P.conv_tc_all[i + 1];
            __nv_bfloat16* ph = to_planes ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + P.actp[i]) : nullptr;
            __nv_bfloat16* pl = to_planes ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + P.actp[i] + P.act_plane_bytes[i]) : nullptr;
            float* of = to_planes ? nullptr : wsf(ws, P.act[i]);
            const int C4 = P.C / 4;
            const long n_pix = n_vec / C4;
            const bool fast = (g.p == 5 || g.p == 2) && 256 % C4 == 0 && n_pix < (1L << 31);
            const int prow = 256 / C4;
            const int nb = (int)std::min<long>((n_pix + prow - 1) / prow, 148L * 16);
            if (fast && g.p == 5) launch_k(bn_relu_pool_fwd_t_kernel<5>, nb, 256, 0, st, y, stat, of, ph, pl, (unsigned)n_pix, g);
            else if (fast) launch_k(bn_relu_pool_fwd_t_kernel<2>, nb, 256, 0, st, y, stat, of, ph, pl, (unsigned)n_pix, g);
            else launch_k(bn_relu_pool_fwd_kernel, ew_blocks(n_vec), 256, 0, st, y, stat, of, ph, pl, n_vec, g);
        }
        SED_POST_LAUNCH();
}
    }

    // ---- BiGRU stack
    const int BT = (int)(B * P.T);
    const float* seq = wsf(ws, P.act[P.n_conv - 1]);
    rc = enqueue_weight_planes();
    if (rc) return rc;
    wp_join.now();
    for (int l = 0; l < P.n_gru; ++l) {
        const int h = P.gh[l], in = P.gin[l];
        float* gi = wsf(ws, P.gi[l]);
{ char _nm[40]; snprintf(_nm, sizeof _nm, "gru%d.proj", l); SED_PROF(_nm, st);
        if (P.gru_tc[l]) {
            // tcgen05: X planes are kept for the backward dW_ih; the W_ih planes were built on the side stream
            const size_t xpb = ((size_t)BT * in * 2 + 1023) & ~(size_t)1023, wpb = ((size_t)6 * h * in * 2 + 1023) & ~(size_t)1023;
            char* xp = reinterpret_cast<char*>(ws) + P.gxp[l];
            const char* wp = reinterpret_cast<const char*>(ws) + P.wihp[l];
            rc = split_planes(seq, xp, xp + xpb, (long)BT * in, st);
            if (rc) return rc;
            rc = gemm_tc(xp, xp + xpb, 0, wp, wp + wpb, 0, BT, 6 * h, in, params + P.bih[l], gi, 6L * h, 0, nullptr, st);
        } else {
            rc = gemm_simt(BT, 6 * h, in, 1, RowMajor{seq, in}, RowMajor{params + P.wih[l], in},
                           EpiStore{gi, 6L * h, params + P.bih[l], 0}, st);
        }
        if (rc) return rc;
}
{ char _nm[40]; snprintf(_nm, sizeof _nm, "gru%d.scan_fwd", l); SED_PROF(_nm, st);
        char* hpl = reinterpret_cast<char*>(ws) + P.hpp[l];
        rc = gru_scan_forward(gi, params + P.whh[l], params + P.bhh[l], wsf(ws, P.gout[l]), wsf(ws, P.gates[l]),
                              batch, P.T, h, st, (training && P.gru_planes[l]) ? hpl : nullptr,
                              (training && P.gru_planes[l]) ? hpl + P.hp_plane_bytes[l] : nullptr);
        if (rc) return rc;
}
        seq = wsf(ws, P.gout[l]);
    }

    // ---- per-frame dense head (skipped when the caller runs the fused head, sedb200_crnn_head_fwd_bwd)
    for (int j = 0; logits && j < P.n_dense; ++j) {
        const bool last = (j == P.n_dense - 1);
        float* out = last ? logits : wsf(ws, P.hid[j]);
{ char _nm[40]; snprintf(_nm, sizeof _nm, "dense%d.fwd", j); SED_PROF(_nm, st);
        rc = gemm_simt(BT, P.dout[j], P.din[j], 1, RowMajor{seq, P.din[j]}, RowMajor{params + P.dn_w[j], P.din[j]},
                       EpiStore{out, P.dout[j], params + P.dn_b[j], (!last && d->dense_relu) ? 1 : 0}, st);
        if (rc) return rc;
}
        seq = out;
    }
    return SEDB200_OK;
}

int sedb200_crnn_head_supported(const sedb200_crnn_desc* d) {
    Plan P;
    if (make_plan(d, 0, &P)) return 0;
    return head_fused_supported(P) ? 1 : 0;
}

int sedb200_crnn_head_fwd_bwd(const sedb200_crnn_desc* d, const float* params, int batch, void* ws, size_t ws_bytes,
                              const float* targets, int loss_kind, float alpha, float gamma, float grad_scale,
                              float* logits, float* probs, float* loss, float* grads, void* stream) {
    Plan P;
    int rc = make_plan(d, batch, &P);
    if (rc) return rc;
    SED_REQUIRE(batch >= 1 && params && ws && targets && loss && grads, SEDB200_EINVAL, "crnn_head_fwd_bwd: bad argument");
    SED_REQUIRE(loss_kind == SEDB200_LOSS_BCE || loss_kind == SEDB200_LOSS_FOCAL, SEDB200_EINVAL, "crnn_head_fwd_bwd: loss kind %d", loss_kind);
    SED_REQUIRE(ws_bytes >= P.ws_bytes, SEDB200_EWORKSPACE, "crnn_head_fwd_bwd: workspace %zu < %zu bytes", ws_bytes, P.ws_bytes);
    SED_REQUIRE(head_fused_supported(P), SEDB200_ESHAPE, "crnn_head_fwd_bwd: needs exactly two small dense layers");
    rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    SED_CUDA_OK(cudaMemsetAsync(grads, 0, (size_t)P.n_params * 4, st));
    return head_fused_run(P, d, params, batch, wsf(ws, P.gout[P.n_gru - 1]), targets, loss_kind, alpha, gamma, grad_scale,
                          logits, probs, loss, wsf(ws, P.dseq[0]), grads, wsf(ws, P.part), st);
}

static int crnn_backward_impl(const sedb200_crnn_desc* d, const float* params, const float* x, int batch,
                              unsigned long long seed, const unsigned long long* seed_ptr, void* ws, size_t ws_bytes,
                              const float* dlogits, float* grads, float* dx, void* stream) {
    Plan P;
    int rc = make_plan(d, batch, &P);
    if (rc) return rc;
    PdlCaptureScope pdl_scope(plan_launch_bound(P));
    SED_REQUIRE(batch >= 1, SEDB200_EINVAL, "crnn_backward: batch %d", batch);
    SED_REQUIRE(params && x && ws && grads, SEDB200_EINVAL, "crnn_backward: null buffer");
    SED_REQUIRE(ws_bytes >= P.ws_bytes, SEDB200_EWORKSPACE, "crnn_backward: workspace %zu < %zu bytes", ws_bytes, P.ws_bytes);
    rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const long B = batch;
    const int BT = (int)(B * P.T);
    float* part = wsf(ws, P.part);
    const int kSplit = 48;

    // zero the alignment padding of the gradient buffer once (tensors themselves are fully overwritten); with
    // dlogits == NULL the fused head has done that, written the dense gradients and left d(gru output) in dseq[0]
    if (dlogits) SED_CUDA_OK(cudaMemsetAsync(grads, 0, (size_t)P.n_params * 4, st));

    // ---- dense head
    const float* dout = dlogits;
    for (int j = P.n_dense - 1; dlogits && j >= 0; --j) {
        char _nm[40]; snprintf(_nm, sizeof _nm, "dense%d.bwd", j); SED_PROF(_nm, st);
        const float* in = j == 0 ? wsf(ws, P.gout[P.n_gru - 1]) : wsf(ws, P.hid[j - 1]);
        const int N = P.dout[j], D = P.din[j];
        float* din = j == 0 ? wsf(ws, P.dseq[0]) : wsf(ws, P.dhid[(j - 1) & 1]);
        const float* mask = (j > 0 && d->dense_relu) ? wsf(ws, P.hid[j - 1]) : nullptr;
        if (dense_small_ok(N, D)) {
            rc = ensure_dyn_smem((const void*)dense_bwd_small_kernel, 160 * 1024);
            if (rc) return rc;
            const int nblk = (BT + kDbRows - 1) / kDbRows;
            launch_k(dense_bwd_small_kernel, nblk, 256, dense_small_smem(N, D), st, dout, in, params + P.dn_w[j], mask, BT, N, D, din, part);
            SED_POST_LAUNCH();
            launch_k(dense_bwd_reduce_kernel, ((N * D + N) * 32 + 255) / 256, 256, 0, st, part, nblk, N * D, N, grads + P.dn_w[j], grads + P.dn_b[j]);
            SED_POST_LAUNCH();
        } else {
        // dW[n][k] = sum_m dout[m][n] * in[m][k]
        const int sp = gemm_simt_splits(BT, kSplit);
        rc = gemm_simt(N, D, BT, kSplit, ColMajor{dout, N}, ColMajor{in, D}, EpiPartial{part, (long)N * D, D}, st);
        if (rc) return rc;
        rc = reduce_partials(part, grads + P.dn_w[j], (long)N * D, sp, st);
        if (rc) return rc;
        rc = colsum(dout, BT, N, grads + P.dn_b[j], part, st);
        if (rc) return rc;
        // d(in)[m][k] = sum_n dout[m][n] * W[n][k]   (through the previous layer's ReLU if any)
        if (mask)
            rc = gemm_simt(BT, D, N, 1, RowMajor{dout, N}, ColMajor{params + P.dn_w[j], D},
                           EpiReluMask{din, D, mask}, st);
        else
            rc = gemm_simt(BT, D, N, 1, RowMajor{dout, N}, ColMajor{params + P.dn_w[j], D},
                           EpiStore{din, D, nullptr, 0}, st);
        if (rc) return rc;
        }
        dout = din;
    }

    // ---- BiGRU stack (dout = grad wrt the layer's output, in dseq[cur])
    int cur = 0;
    SideStream* side = nullptr;
    bool side_busy[4] = {false, false, false, false};
    for (int l = P.n_gru - 1; l >= 0; --l) {
        const int h = P.gh[l], in = P.gin[l];
        const float* xin = l == 0 ? wsf(ws, P.act[P.n_conv - 1]) : wsf(ws, P.gout[l - 1]);
        float* dgi = wsf(ws, P.dgi);
        float* dgh = wsf(ws, P.dgh);
        const bool fusedg = gru_scan_fused_param_grads(h);
        // two-stream layer: the scan emits the tensor-core operand planes itself, so everything that is NOT on the path to
        // d(input) (bias reduction, dW_hh, dW_ih) can run beside the main stream
        const bool two_stream = P.gru_planes[l] && P.gru_tc[l] && fusedg && gemm_tc_supported(6 * h, 2 * h, BT) && P.dgp2 != 0;
        const int par = l & 1;
        float* part_w = part;                                      // [B][2][3h][h]
        float* part_b = two_stream ? wsf(ws, P.gbias[par]) : part + (size_t)B * 6 * h * h;   // [B][2][2][3h]
        char* dgpl = reinterpret_cast<char*>(ws) + ((two_stream && par) ? P.dgp2 : P.dgp);   // {dgi_hi, dgi_lo, dgh_hi, dgh_lo} when the scan emits planes
        void* planes[4] = {dgpl, dgpl + P.dg_plane_bytes, dgpl + 2 * P.dg_plane_bytes, dgpl + 3 * P.dg_plane_bytes};
        if (two_stream) {
            if (!side) { rc = side_stream(&side); if (rc) return rc; }
            // an earlier layer of the same parity may still be reading these planes on the side stream
            if (side_busy[par]) SED_CUDA_OK(cudaStreamWaitEvent(st, side->done[par], 0));
        }
{ char _nm[40]; snprintf(_nm, sizeof _nm, "gru%d.scan_bwd", l); SED_PROF(_nm, st);
        rc = gru_scan_backward(wsf(ws, P.dseq[cur]), wsf(ws, P.gout[l]), wsf(ws, P.gates[l]), params + P.whh[l],
                               dgi, dgh, part_w, part_b, batch, P.T, h, st, P.gru_planes[l] ? planes : nullptr);
        if (rc) return rc;
}
        if (two_stream) {
            char _nm3[40]; snprintf(_nm3, sizeof _nm3, "gru%d.bwd_dx", l); SED_PROF(_nm3, st);
            cudaStream_t ss = side->st;
            SED_CUDA_OK(cudaEventRecord(side->fork[par], st));
            SED_CUDA_OK(cudaStreamWaitEvent(ss, side->fork[par], 0));
            // ---- side stream: bias gradients, dW_hh, dW_ih
            rc = reduce_bias_partials(part_b, grads + P.bih[l], grads + P.bhh[l], 6 * h, batch, ss);
            if (rc) return rc;
            const char *gi_hi = dgpl, *gi_lo = dgpl + P.dg_plane_bytes;
            const char *gh_hi = dgpl + 2 * P.dg_plane_bytes, *gh_lo = dgpl + 3 * P.dg_plane_bytes;
            const char* hp_hi = reinterpret_cast<const char*>(ws) + P.hpp[l];
            const char* hp_lo = hp_hi + P.hp_plane_bytes[l];
            float* stmp = wsf(ws, P.tc_side);
            float* stpart = stmp + 12 * h * h + 64;
            rc = gemm_tc(gh_hi, gh_lo, 1, hp_hi, hp_lo, 1, 6 * h, 2 * h, BT, nullptr, stmp, 2 * h, 1, stpart, ss);
            if (rc) return rc;
            launch_k(extract_whh_kernel, (6 * h * h + 255) / 256, 256, 0, ss, stmp, h, grads + P.whh[l]);
            SED_POST_LAUNCH();
            const size_t xpb = ((size_t)BT * in * 2 + 1023) & ~(size_t)1023, wpb = ((size_t)6 * h * in * 2 + 1023) & ~(size_t)1023;
            const char* xp = reinterpret_cast<const char*>(ws) + P.gxp[l];
            // dW_ih[n6][k] = sum_m dgi[m][n6] * xin[m][k]        (both operands stored [B*T][.]: MN-major)
            rc = gemm_tc(gi_hi, gi_lo, 1, xp, xp + xpb, 1, 6 * h, in, BT, nullptr, grads + P.wih[l], in, 1, wsf(ws, P.tc_side), ss);
            if (rc) return rc;
            SED_CUDA_OK(cudaEventRecord(side->done[par], ss));
            side_busy[par] = true;
            // ---- main stream: d(xin)[m][k] = sum_n6 dgi[m][n6] * W_ih[n6][k]    (W_ih stored [n6][k] = [K][N]: MN-major)
            const char* wp = reinterpret_cast<const char*>(ws) + P.wihp[l];          // built by the forward pass
            float* dxin2 = wsf(ws, P.dseq[cur ^ 1]);
            rc = gemm_tc(gi_hi, gi_lo, 0, wp, wp + wpb, 1, BT, in, 6 * h, nullptr, dxin2, in, 0, nullptr, st);
            if (rc) return rc;
            cur ^= 1;
            continue;
        }
        char _nm2[40]; snprintf(_nm2, sizeof _nm2, "gru%d.bwd_gemms", l); SED_PROF(_nm2, st);
        if (fusedg) {
            // the scan left per-batch-row partials of both bias gradients: fixed-order sum over B
            rc = reduce_bias_partials(part_b, grads + P.bih[l], grads + P.bhh[l], 6 * h, batch, st);
            if (rc) return rc;
        } else {
            rc = colsum(dgi, BT, 6 * h, grads + P.bih[l], part, st);
            if (rc) return rc;
            rc = colsum(dgh, BT, 6 * h, grads + P.bhh[l], part, st);
            if (rc) return rc;
        }
        if (P.gru_tc[l] && gemm_tc_supported(6 * h, 2 * h, BT)) {
            // dW_hh on tcgen05: [6h][2h] = dgh^T (B*T x 6h) . h_prev (B*T x 2h); the two diagonal blocks are
            // the per-direction gradients (the off-diagonal blocks are discarded)
            const size_t gpb = ((size_t)BT * 6 * h * 2 + 1023) & ~(size_t)1023, hpb = ((size_t)BT * 2 * h * 2 + 1023) & ~(size_t)1023;
            char* gp = reinterpret_cast<char*>(ws) + P.tc;
            char* hp = gp + 2 * gpb;
            float* tmp = reinterpret_cast<float*>(hp + 2 * hpb);
            float* tpart = tmp + 12 * h * h + 64;
            const char *gh_hi = gp, *gh_lo = gp + gpb, *hp_hi = hp, *hp_lo = hp + hpb;
            if (P.gru_planes[l]) {                          // both operands were written by the scans themselves
                gh_hi = dgpl + 2 * P.dg_plane_bytes; gh_lo = dgpl + 3 * P.dg_plane_bytes;
                hp_hi = reinterpret_cast<char*>(ws) + P.hpp[l]; hp_lo = hp_hi + P.hp_plane_bytes[l];
            } else {
                rc = split_planes(dgh, gp, gp + gpb, (long)BT * 6 * h, st);
                if (rc) return rc;
                rc = hprev_planes(wsf(ws, P.gout[l]), hp, hp + hpb, BT, P.T, h, st);
                if (rc) return rc;
            }
            rc = gemm_tc(gh_hi, gh_lo, 1, hp_hi, hp_lo, 1, 6 * h, 2 * h, BT, nullptr, tmp, 2 * h, 1, tpart, st);
            if (rc) return rc;
            launch_k(extract_whh_kernel, (6 * h * h + 255) / 256, 256, 0, st, tmp, h, grads + P.whh[l]);
            SED_POST_LAUNCH();
        } else {
            for (int dir = 0; dir < 2; ++dir) {
                const int spw = gemm_simt_splits(BT, kSplit);
                rc = gemm_simt(3 * h, h, BT, kSplit, ColMajor{dgh + dir * 3 * h, 6L * h},
                               HPrevB{wsf(ws, P.gout[l]), P.T, h, dir}, EpiPartial{part, 3L * h * h, h}, st);
                if (rc) return rc;
                rc = reduce_partials(part, grads + P.whh[l] + (long)dir * 3 * h * h, 3L * h * h, spw, st);
                if (rc) return rc;
            }
        }
        float* dxin = wsf(ws, P.dseq[cur ^ 1]);
        if (P.gru_tc[l]) {
            const size_t xpb = ((size_t)BT * in * 2 + 1023) & ~(size_t)1023, wpb = ((size_t)6 * h * in * 2 + 1023) & ~(size_t)1023;
            const size_t gpb = ((size_t)BT * 6 * h * 2 + 1023) & ~(size_t)1023;
            char* xp = reinterpret_cast<char*>(ws) + P.gxp[l];
            char* gp = reinterpret_cast<char*>(ws) + P.tc;
            char* wp = gp + 2 * gpb;
            float* tpart = reinterpret_cast<float*>(wp + 2 * wpb);
            const char *gi_hi = gp, *gi_lo = gp + gpb;
            if (P.gru_planes[l]) {
                gi_hi = dgpl; gi_lo = dgpl + P.dg_plane_bytes;
            } else {
                rc = split_planes(dgi, gp, gp + gpb, (long)BT * 6 * h, st);
                if (rc) return rc;
            }
            const char* wihp = reinterpret_cast<const char*>(ws) + P.wihp[l];          // built by the forward pass
            // dW_ih[n6][k] = sum_m dgi[m][n6] * xin[m][k]        (both operands stored [B*T][.]: MN-major)
            rc = gemm_tc(gi_hi, gi_lo, 1, xp, xp + xpb, 1, 6 * h, in, BT, nullptr, grads + P.wih[l], in, 1, tpart, st);
            if (rc) return rc;
            // d(xin)[m][k] = sum_n6 dgi[m][n6] * W_ih[n6][k]    (W_ih stored [n6][k] = [K][N]: MN-major)
            rc = gemm_tc(gi_hi, gi_lo, 0, wihp, wihp + wpb, 1, BT, in, 6 * h, nullptr, dxin, in, 0, nullptr, st);
            if (rc) return rc;
        } else {
        // dW_ih[n6][k] = sum_m dgi[m][n6] * xin[m][k]
        int sp = gemm_simt_splits(BT, kSplit);
        rc = gemm_simt(6 * h, in, BT, kSplit, ColMajor{dgi, 6L * h}, ColMajor{xin, in},
                       EpiPartial{part, 6L * h * in, in}, st);
        if (rc) return rc;
        rc = reduce_partials(part, grads + P.wih[l], 6L * h * in, sp, st);
        if (rc) return rc;
        rc = gemm_simt(BT, in, 6 * h, 1, RowMajor{dgi, 6L * h}, ColMajor{params + P.wih[l], in},
                       EpiStore{dxin, in, nullptr, 0}, st);
        if (rc) return rc;
        }
        cur ^= 1;
    }

    // ---- conv blocks (dA = grad wrt block output; for the last block it is dseq[cur] in [B][T][flat])
    const float* dA = wsf(ws, P.dseq[cur]);
    struct SideJoin {           // the weight-gradient GEMMs rejoin the caller's stream on every way out of this function
        SideStream* s; const bool* busy; cudaStream_t st;
        ~SideJoin() { for (int i = 0; s && i < 4; ++i) if (busy[i]) cudaStreamWaitEvent(st, s->done[i], 0); }
    } side_join{side, side_busy, st};
    // Weight gradients of the plane-native blocks beside the main stream: wgrad_tc_kernel (TMA / tensor work, 0.07-0.12 ms
    // at C2) starts when the block's data gradient has finished and then shares the SMs with the HBM-bound BatchNorm /
    // pool backward passes of the next block (or with block 0's backward): nothing downstream reads dW before the
    // optimizer.  The dy planes and their scale pair alternate by block parity so that the next block's passes do not
    // overwrite what the side stream still reads.  Phase profiling keeps everything on one stream (serial phases).
    // SEDB200_WGRAD_SIDE: 0 = off (one stream), n >= 2 = on with n pipeline stages in the weight-gradient kernel
    // (default 3: C2 step 1.454 ms on one stream, 1.447 / 1.400 / 1.406 ms with 4 / 3 / 2 stages)
    static const int wgrad_side_stages = [] { const char* e = std::getenv("SEDB200_WGRAD_SIDE"); return e ? std::atoi(e) : 3; }();
    const bool wgrad_side = wgrad_side_stages >= 2 && !prof_on();
    for (int i = P.n_conv - 1; i >= 0; --i) {
        const PoolGeom g = pool_geom(P, d, i, 1, seed, seed_ptr);
        const float* y = wsf(ws, P.y[i]);
        const float* stat = wsf(ws, P.stat[i]);
        if (i == 0 && conv0_lean_ok(P.cin[0], P.C, P.pool[0], P.n_conv)) {
            const int gpi = (P.H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch;
            const size_t sm = 2 * (size_t)P.cin[0] * (kC0Rows + 2) * (P.win[0] + 2) * 4;
            const int cin0 = P.cin[0];
            if (!dx) {
                // lean block 0: winners' contributions from dA + winner bytes + input rows, the dense terms from
                // the patch moments the forward pass left in the workspace
                SED_PROF("conv0.bwd_lean", st);
                rc = conv0_lean_backward(x, cin0, batch,
                                         reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(ws) + P.arg0), dA, g,
                                         reinterpret_cast<const double*>(reinterpret_cast<const char*>(ws) + P.gram),
                                         params + P.conv_w[0], params + P.conv_b[0], params + P.bn_w[0], stat, part,
                                         grads + P.conv_w[0], grads + P.conv_b[0], grads + P.bn_w[0], grads + P.bn_b[0], st);
                if (rc) return rc;
                break;
            }
            // the caller wants d(input): rebuild y0 (bit-identical to what the fused forward consumed) and take the
            // general route below
            SED_PROF("conv0.recompute_y", st);
            const dim3 grid(std::min(n_groups, 2 * sm_count()), P.C / 128);
            if (cin0 == 1)
                launch_k(conv0_fwd_stats_kernel<1>, grid, 256, sm, st, x, params + P.conv_w[0], params + P.conv_b[0], wsf(ws, P.y[0]), P.H, P.win[0], P.C, gpi, n_groups, part);
            else
                launch_k(conv0_fwd_stats_kernel<2>, grid, 256, sm, st, x, params + P.conv_w[0], params + P.conv_b[0], wsf(ws, P.y[0]), P.H, P.win[0], P.C, gpi, n_groups, part);
            SED_POST_LAUNCH();
        }
        const long n_pix_out = B * P.H * P.wout[i];
        const long n_elem = B * P.H * P.win[i];                 // BN population per channel
        const int rows = 256 / (P.C / 4);
        const int nblk = (int)std::min<long>((n_pix_out + rows - 1) / rows, 148L * 16);
        int sblk = nblk;                                          // blocks (= partials) of the sums pass
        bool have_amax = false;                                   // the sums pass left max |dz| per block (fp16 dy plane)
{ char _nm[40]; snprintf(_nm, sizeof _nm, "pool%d.bwd_sums", i); SED_PROF(_nm, st);
        const bool idx32 = n_pix_out < (1L << 31);
        const bool planes_out = (i + 1 < P.n_conv) && P.conv_tc_all[i + 1];      // what the forward pass stored
        if (idx32 && 256 % (P.C / 4) == 0) {
            const char* ap = reinterpret_cast<const char*>(ws) + P.actp[i];
            // few long-lived blocks (4 per SM): the per-block prologue / shared-memory reduction / partial store and
            // the finalizer's work shrink with the block count, the loads in flight do not
            sblk = (int)std::min<long>(nblk, 148L * 4);
            launch_k(bn_bwd_sums_act_kernel, sblk, 256, 0, st,
                planes_out ? nullptr : wsf(ws, P.act[i]),
                planes_out ? reinterpret_cast<const __nv_bfloat16*>(ap) : nullptr,
                planes_out ? reinterpret_cast<const __nv_bfloat16*>(ap + P.act_plane_bytes[i]) : nullptr, stat, dA,
                (unsigned)n_pix_out, g, part, P.conv_tc_all[i] ? wsf(ws, P.dys) + 8 : nullptr);
            have_amax = P.conv_tc_all[i];
        } else if (g.p == 5 && idx32) launch_k(bn_pool_bwd_sums_t_kernel<5>, nblk, 256, 0, st, y, stat, dA, n_pix_out, g, part);
        else if (g.p == 2 && idx32) launch_k(bn_pool_bwd_sums_t_kernel<2>, nblk, 256, 0, st, y, stat, dA, n_pix_out, g, part);
        else launch_k(bn_pool_bwd_sums_kernel, nblk, 256, 0, st, y, stat, dA, n_pix_out, g, part);
        SED_POST_LAUNCH();
}
        float* bnsum = wsf(ws, P.bnsum);
        launch_k(bn_bwd_finalize_kernel, P.C, 128, 0, st, part, sblk, P.C, n_elem, grads + P.bn_w[i],
                                                                 grads + P.bn_b[i], bnsum);
        SED_POST_LAUNCH();
        if (i == 0 && !dx && conv0_direct_ok(P.cin[0], P.C)) {
            // fused BN/ReLU/pool backward + weight/bias gradient: dy0 is never materialised
            SED_PROF("conv0.bwd_fused", st);
            const dim3 grid((P.H + kC0Rows - 1) / kC0Rows, batch, P.C / 128);
            const size_t sm = (size_t)P.cin[0] * (kC0Rows + 2) * (P.win[0] + 2) * 4;
            const int J = P.cin[0] * 9;
            const bool exact = (g.W == g.Wo * g.p) && (g.p == 5 || g.p == 2);
            int nparts = (int)(grid.x * grid.y);
            if (exact) {
                const int gpi = (P.H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch;
                const dim3 pgrid(std::min(n_groups, sm_count()), P.C / 128);
                nparts = (int)pgrid.x;
                if (g.p == 5 && P.cin[0] == 1) launch_k(conv0_bwd_fused_t_kernel<1, 5>, pgrid, 256, 2 * sm, st, x, y, stat, dA, bnsum, g, gpi, n_groups, part);
                else if (g.p == 5) launch_k(conv0_bwd_fused_t_kernel<2, 5>, pgrid, 256, 2 * sm, st, x, y, stat, dA, bnsum, g, gpi, n_groups, part);
                else if (P.cin[0] == 1) launch_k(conv0_bwd_fused_t_kernel<1, 2>, pgrid, 256, 2 * sm, st, x, y, stat, dA, bnsum, g, gpi, n_groups, part);
                else launch_k(conv0_bwd_fused_t_kernel<2, 2>, pgrid, 256, 2 * sm, st, x, y, stat, dA, bnsum, g, gpi, n_groups, part);
            } else if (P.cin[0] == 1)
                launch_k(conv0_bwd_fused_kernel<1>, grid, 256, sm, st, x, y, stat, dA, bnsum, g, part);
            else
                launch_k(conv0_bwd_fused_kernel<2>, grid, 256, sm, st, x, y, stat, dA, bnsum, g, part);
            SED_POST_LAUNCH();
            launch_k(conv0_bwd_reduce_kernel, ((J + 1) * P.C * 32 + 255) / 256, 256, 0, st,
                part, nparts, J, P.C, grads + P.conv_w[0], grads + P.conv_b[0]);
            SED_POST_LAUNCH();
            break;
        }
        float* dy = wsf(ws, P.dy);
        const long n_vec = n_pix_out * (P.C / 4);
        // plane-native block: dy exists only as ONE fp16 plane of dy * scale (crnn_block.cuh: store_dy4); the scale
        // {s, 1/s} lives in the workspace and is read by the kernels that produce / consume the plane
        const int cpar = i & 1;
        __nv_bfloat16* dyh = P.conv_tc_all[i] ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + P.dyp + cpar * P.dy_plane_bytes) : nullptr;
        float* dys = wsf(ws, P.dys) + 2 * cpar;                    // {s, 1/s} of this block's dy plane
        if (P.conv_tc_all[i]) {
            // an earlier block of the same parity may still be read by its weight-gradient kernel on the side stream
            if (side_busy[2 + cpar]) { SED_CUDA_OK(cudaStreamWaitEvent(st, side->done[2 + cpar], 0)); side_busy[2 + cpar] = false; }
            if (have_amax) {
                launch_k(dy_scale_kernel, 1, 128, 0, st, wsf(ws, P.dys) + 8, sblk, stat, bnsum, P.C, dys);
                SED_POST_LAUNCH();
            } else {
                return fail(SEDB200_ESHAPE, "crnn_backward: plane-native block %d without the activation-sums pass", i);
            }
        }
{ char _nm[40]; snprintf(_nm, sizeof _nm, "pool%d.bwd_dy", i); SED_PROF(_nm, st);
        float* dyf = P.conv_tc_all[i] ? nullptr : dy;
        const bool exact_t = g.W == g.Wo * g.p && (g.p == 5 || g.p == 2) && 256 % (P.C / 4) == 0 && n_pix_out < (1L << 31);
        if (exact_t && g.p == 5)
            launch_k(bn_pool_bwd_dy_t_kernel<5>, nblk, 256, 0, st, y, stat, dA, bnsum, (unsigned)n_pix_out, g, dyf, dyh, dyh ? dys : nullptr);
        else if (exact_t)
            launch_k(bn_pool_bwd_dy_t_kernel<2>, nblk, 256, 0, st, y, stat, dA, bnsum, (unsigned)n_pix_out, g, dyf, dyh, dyh ? dys : nullptr);
        else
            launch_k(bn_pool_bwd_dy_kernel, ew_blocks(n_vec), 256, 0, st, y, stat, dA, bnsum, n_vec, g, dyf, dyh, dyh ? dys : nullptr);
        SED_POST_LAUNCH();
}

        const float* in = i == 0 ? x : wsf(ws, P.act[i - 1]);
        const InStrides s = in_strides(P, d, i);
        const int M = (int)n_elem, J = P.cin[i] * 9;
        // conv bias grad = column sums of dy.  Through train-mode BatchNorm this is identically zero
        // (sum_p dy = scale * (sum dz - n*mean(dz) - mean(dz*xhat) * sum xhat) = 0); what PyTorch stores is the
        // fp32 rounding residue of that sum (~1e-10).  grads was zero-filled above: the exact value stays.
        // wgrad
        const int want = std::max(1, std::min(64, M / 2048));
        const int sp = gemm_simt_splits(M, want);
        const bool wg_side = wgrad_side && P.conv_tc_all[i] && i > 0;
{ char _nm[40]; snprintf(_nm, sizeof _nm, "conv%d.wgrad", i); SED_PROF(_nm, st);
        if (wg_side) {
            // after the data gradient, below
        } else if (P.conv_tc_all[i]) {
            const char* xp = reinterpret_cast<const char*>(ws) + P.actp[i - 1];
            float* wpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + P.tc + conv_tc_weight_scratch_bytes(P.cin[i], P.C));
            rc = wgrad_tc_planes(dyh, nullptr, xp, nullptr, grads + P.conv_w[i], batch, P.H, P.win[i], P.cin[i], P.C, wpart,
                                 st, 1, kPlaneF16, dys + 1);
            if (rc) return rc;
        } else if (i > 0 && d->tensor_cores && wgrad_tc_supported(P.H, P.win[i], P.cin[i], P.C)) {
            rc = wgrad_tc(dy, in, grads + P.conv_w[i], batch, P.H, P.win[i], P.cin[i], P.C, wsf(ws, P.tc), P.tc_bytes, st);
            if (rc) return rc;
        } else if (conv_small_wgrad_supported(P.cin[i], P.C, P.win[i]) &&
                   conv_small_wgrad_part_floats(P.cin[i], P.C, batch, P.H) <= P.part_floats) {
            rc = conv_small_wgrad(dy, in, s.sB, s.sH, s.sW, s.sC, P.cin[i], P.C, batch, P.H, P.win[i], part,
                                  grads + P.conv_w[i], st);
            if (rc) return rc;
        } else {
        rc = gemm_simt(P.C, J, M, want, ColMajor{dy, P.C}, ConvWgradB{in, P.H, P.win[i], s.sB, s.sH, s.sW, s.sC},
                       EpiPartial{part, (long)P.C * J, J}, st);
        if (rc) return rc;
        rc = reduce_partials(part, grads + P.conv_w[i], (long)P.C * J, sp, st);
        if (rc) return rc;
        }
}
        // dgrad
        if (i > 0) {
            float* dprev = wsf(ws, P.dact[i & 1]);
{ char _nm[40]; snprintf(_nm, sizeof _nm, "conv%d.dgrad", i); SED_PROF(_nm, st);
            if (P.conv_tc_all[i])
                rc = conv_tc_planes_w(dyh, nullptr, reinterpret_cast<const char*>(ws) + P.wpl[i][1], nullptr, dprev, nullptr,
                                      batch, P.H, P.win[i], P.cin[i], P.C, 1, st, 1, kPlaneF16, dys + 1);   // weight planes built by the forward pass
            else if (d->tensor_cores && conv_tc_supported(P.H, P.win[i], P.C, P.cin[i]))
                rc = conv_tc_forward(dy, params + P.conv_w[i], nullptr, dprev, batch, P.H, P.win[i], P.cin[i], P.C, 1,
                                     wsf(ws, P.tc), P.tc_bytes, st);
            else if (conv_small_supported(P.C, P.cin[i], P.win[i]))
                rc = conv_small_forward(dy, (long)P.H * P.win[i] * P.C, (long)P.win[i] * P.C, P.C, 1, P.C, batch, P.H,
                                        P.win[i], params + P.conv_w[i], nullptr, P.cin[i], 1, dprev, st);
            else
                rc = gemm_simt(M, P.cin[i], 9 * P.C, 1, ConvDgradA{dy, P.H, P.win[i], P.C},
                               ConvDgradB{params + P.conv_w[i], P.cin[i], P.C}, EpiStore{dprev, P.cin[i], nullptr, 0}, st);
            if (rc) return rc;
}
            if (wg_side) {
                if (!side) { rc = side_stream(&side); if (rc) return rc; side_join.s = side; }
                cudaStream_t ss = side->st;
                SED_CUDA_OK(cudaEventRecord(side->fork[2 + cpar], st));
                SED_CUDA_OK(cudaStreamWaitEvent(ss, side->fork[2 + cpar], 0));
                const char* xp = reinterpret_cast<const char*>(ws) + P.actp[i - 1];
                float* wpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + P.tc + conv_tc_weight_scratch_bytes(P.cin[i], P.C));
                // three pipeline stages instead of six: 97 KB of shared memory, so that the kernels of the main stream
                // (block 0's backward: 95 KB per CTA) fit on the SM beside it
                rc = wgrad_tc_planes(dyh, nullptr, xp, nullptr, grads + P.conv_w[i], batch, P.H, P.win[i], P.cin[i], P.C, wpart,
                                     ss, 1, kPlaneF16, dys + 1, wgrad_side_stages);
                if (rc) return rc;
                SED_CUDA_OK(cudaEventRecord(side->done[2 + cpar], ss));
                side_busy[2 + cpar] = true;
            }
            dA = dprev;
        } else if (dx) {
            rc = gemm_simt(M, P.cin[0], 9 * P.C, 1, ConvDgradA{dy, P.H, P.win[0], P.C},
                           ConvDgradB{params + P.conv_w[0], P.cin[0], P.C},
                           EpiStrided{dx, P.H, P.win[0], s.sB, s.sH, s.sW, s.sC}, st);
            if (rc) return rc;
        }
    }
    return SEDB200_OK;
}

int sedb200_crnn_forward(const sedb200_crnn_desc* d, const float* params, float* bn_state, const float* x,
                         int batch, int training, unsigned long long seed, void* ws, size_t ws_bytes,
                         float* logits, void* stream) {
    return crnn_forward_impl(d, params, bn_state, x, batch, training, seed, nullptr, ws, ws_bytes, logits, stream);
}
int sedb200_crnn_forward_s(const sedb200_crnn_desc* d, const float* params, float* bn_state, const float* x,
                           int batch, int training, const void* step_state, void* ws, size_t ws_bytes,
                           float* logits, void* stream) {
    SED_REQUIRE(step_state, SEDB200_EINVAL, "crnn_forward_s: null step state");
    return crnn_forward_impl(d, params, bn_state, x, batch, training, 0, reinterpret_cast<const unsigned long long*>(step_state),
                             ws, ws_bytes, logits, stream);
}
int sedb200_crnn_backward(const sedb200_crnn_desc* d, const float* params, const float* x, int batch,
                          unsigned long long seed, void* ws, size_t ws_bytes, const float* dlogits,
                          float* grads, float* dx, void* stream) {
    return crnn_backward_impl(d, params, x, batch, seed, nullptr, ws, ws_bytes, dlogits, grads, dx, stream);
}
int sedb200_crnn_backward_s(const sedb200_crnn_desc* d, const float* params, const float* x, int batch,
                            const void* step_state, void* ws, size_t ws_bytes, const float* dlogits,
                            float* grads, float* dx, void* stream) {
    SED_REQUIRE(step_state, SEDB200_EINVAL, "crnn_backward_s: null step state");
    return crnn_backward_impl(d, params, x, batch, 0, reinterpret_cast<const unsigned long long*>(step_state), ws, ws_bytes,
                              dlogits, grads, dx, stream);
}

int sedb200_crnn_dropout_mask(const sedb200_crnn_desc* d, int batch, unsigned long long seed, int block,
                              unsigned char* mask_dev, void* stream) {
    Plan P;
    int rc = make_plan(d, 0, &P);
    if (rc) return rc;
    SED_REQUIRE(batch >= 1 && mask_dev, SEDB200_EINVAL, "crnn_dropout_mask: batch %d / null buffer", batch);
    SED_REQUIRE(block >= 0 && block < P.n_conv, SEDB200_EINVAL, "crnn_dropout_mask: block %d of %d", block, P.n_conv);
    rc = require_sm100();
    if (rc) return rc;
    const PoolGeom g = pool_geom(P, d, block, 1, seed);
    const long n_vec = (long)batch * P.H * P.wout[block] * (P.C / 4);
    launch_k(dropout_mask_kernel, ew_blocks(n_vec), 256, 0, as_stream(stream), mask_dev, n_vec, g);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // extern "C"
