// conv0_lean.cu -- the first conv block without its output tensor ("lean block 0"): patch-moment BatchNorm statistics,
// fused conv + BN + ReLU + max-pool + dropout forward (fp32 CUDA cores or tcgen05), sparse winner backward.
// Reference ops: crnn_lightning.py:46-52 / sed.py:87-92,106-107 (Conv2d -> BatchNorm2d -> ReLU -> MaxPool2d((1,p)) ->
// Dropout) of the FIRST block, and their backward.  Orchestrated by crnn.cu through conv0_lean.cuh.
#include <cuda_bf16.h>
#include "conv0_lean.cuh"
#include "tc_umma.cuh"

#include <algorithm>
#include <cstdlib>

namespace sedb200 {
namespace {

// ----------------------------------------------------------------------------- first conv block, lean
// The output of conv 0 (4 B x B*H*W*C: 671 MB at C2, 10.7 GB at C5) is the largest tensor of the step, and
// with K = 9*Cin <= 18 it is also the cheapest one to describe.  This path never writes or reads it:
//
//   * BatchNorm statistics of y = w.patch + b are a function of the first two moments of the PATCHES:
//         mean_c = b_c + w_c.m,   var_c = w_c^T (G - m m^T) w_c,   m = E[patch], G = E[patch patch^T]  (K x K)
//     so one pass over the INPUT (conv0_gram_kernel) replaces the statistics pass over y.
//   * The forward kernel knows scale / shift before it starts: conv + BN + ReLU + max-pool + dropout happen in
//     registers; it writes the pooled block output and one byte per (window, channel): the winner's position in
//     the window, bit 7 set when the ReLU or the dropout killed the window.
//   * Backward: dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) is dense in the pixels, but the only dense
//     terms are (constant) and (xhat): sum_pix patch_k and sum_pix xhat*patch_k follow from (m, G) again.  What
//     needs the data is S_k = sum over WINDOWS of dz * patch_k(winner) (one pixel in p), and
//     sum dz*xhat(winner) = invstd * (b*sum dz + w.S - mean*sum dz) needs nothing else.  So one kernel reads
//     dA + the winner bytes + the input rows, and a per-channel finalizer closes d(gamma), d(beta) and dW.
//     The conv bias gradient through train-mode BatchNorm is identically zero and is written as 0.
//
// HBM traffic of block 0 at C2: forward |x| + |a| + |a|/4 = 178 MB (was 671 MB written + 805 MB read back),
// backward |dA| + |a|/4 + |x| = 178 MB (was 268 + 816 MB).
// ---- asynchronous, division-free staging of the input rows of one row group (8 rows + halo, zero-filled borders):
// warp <-> (input channel, row), lane <-> column; 4-byte cp.async with src-size 0 for the zero fill
__device__ __forceinline__ void cpa4_zfill(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int CIN>
__device__ __forceinline__ void stage_x_rows_async(const float* __restrict__ x, float* xs, int b, int h0, int H, int W,
                                                   int row_stride = 0) {
    const int Wp = row_stride ? row_stride : W + 2, nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rowid = warp; rowid < CIN * (kC0Rows + 2); rowid += nwarps) {
        const int ci = rowid / (kC0Rows + 2), rr = rowid - ci * (kC0Rows + 2);      // constant divisor: no XU work
        const int hh = h0 - 1 + rr;
        const bool rowok = hh >= 0 && hh < H;
        const float* src = x + (((long)b * CIN + ci) * H + (rowok ? hh : 0)) * W;
        float* dst = xs + rowid * Wp;
        for (int cc = lane; cc < W + 2; cc += 32) {
            const bool ok = rowok && cc >= 1 && cc <= W;
            cpa4_zfill(dst + cc, src + (ok ? cc - 1 : 0), ok);
        }
    }
}

// (image, first row) of the row groups a persistent block walks, advanced without divisions
struct GroupWalk {
    int b, gi, step_b, step_g, gpi;
    __device__ GroupWalk(int grp0, int stride, int groups_per_img) : gpi(groups_per_img) {
        b = grp0 / gpi; gi = grp0 - b * gpi;
        step_b = stride / gpi; step_g = stride - step_b * gpi;
    }
    __device__ void next() {
        b += step_b; gi += step_g;
        if (gi >= gpi) { gi -= gpi; ++b; }
    }
    __device__ int h0() const { return gi * kC0Rows; }
};

template <int CIN>
struct GramDims {
    static constexpr int K = CIN * 9;
    static constexpr int TEAMS = CIN * 3;               // a team = one warp = (input channel, kernel row) of the OWN entry
    static constexpr int KP = (K + 2) / 2;              // column pairs: k' = 0..K-1, then the constant 1 (plain sums)
    static constexpr int THREADS = TEAMS * 32;
};

// part[blk][k][k'] = sum over the block's pixels of patch[k] * patch[k'] (k' = K: the plain sum).  Warp (ci, r) owns
// the three rows k = ci*9 + r*3 + t; lanes walk the pixels of the row group.  Per-thread fp32 sums of ~140 products
// (packed fma.f32x2 over column pairs), fixed shuffle tree, doubles afterwards.  `wmagic` = ceil(2^32 / W):
// row = umulhi(p, wmagic) for p < 2^16.
template <int CIN>
__global__ void __launch_bounds__(GramDims<CIN>::THREADS)
conv0_gram_kernel(const float* __restrict__ x, int H, int W, unsigned wmagic, int groups_per_img, int n_groups,
                  float* __restrict__ part) {
    pdl_wait();
    using D = GramDims<CIN>;
    extern __shared__ float xs_all[];                     // 2 x [CIN][10][W+32]: row stride = W (mod 32), so the
                                                          // lanes that wrap to the next image row keep walking the banks
    const int team = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tci = team / 3, tr = team - tci * 3;
    const int Wp = W + 32, xsz = CIN * (kC0Rows + 2) * Wp;
    const int own_off = (tci * (kC0Rows + 2) + tr) * Wp;
    float2 acc[3][D::KP];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < D::KP; ++k) acc[a][k] = make_float2(0.0f, 0.0f);
    int grp = blockIdx.x, buf = 0;
    GroupWalk gw(grp, gridDim.x, groups_per_img), gn = gw;
    if (grp < n_groups) stage_x_rows_async<CIN>(x, xs_all, gw.b, gw.h0(), H, W, Wp);
    cpa_commit();
    cpa_wait_all();
    __syncthreads();
    for (; grp < n_groups; grp += gridDim.x, buf ^= 1) {
        const float* xs = xs_all + buf * xsz;
        const int h0 = gw.h0();
        gn.next();
        if (grp + (int)gridDim.x < n_groups) stage_x_rows_async<CIN>(x, xs_all + (buf ^ 1) * xsz, gn.b, gn.h0(), H, W, Wp);
        cpa_commit();
        const int npix = min(kC0Rows, H - h0) * W;
        for (int p = lane; p < npix; p += 32) {
            const int row = (int)__umulhi((unsigned)p, wmagic), col = p - row * W;
            const float* base = xs + row * Wp + col;
            float pv[2 * D::KP], own[3];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int t = 0; t < 3; ++t) pv[ci * 9 + r * 3 + t] = base[(ci * (kC0Rows + 2) + r) * Wp + t];
            pv[D::K] = 1.0f;
            if (D::K + 1 < 2 * D::KP) pv[2 * D::KP - 1] = 0.0f;
#pragma unroll
            for (int t = 0; t < 3; ++t) own[t] = base[own_off + t];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float2 o2 = make_float2(own[a], own[a]);
#pragma unroll
                for (int k = 0; k < D::KP; ++k) acc[a][k] = __ffma2_rn(o2, make_float2(pv[2 * k], pv[2 * k + 1]), acc[a][k]);
            }
        }
        gw = gn;
        cpa_wait_all();
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k <= D::K; ++k) {
            float v = (k & 1) ? acc[a][k >> 1].y : acc[a][k >> 1].x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[((long)blockIdx.x * D::K + team * 3 + a) * (D::K + 1) + k] = v;
        }
}

// ----------------------------------------------------------------------------- patch moments from autocorrelations
// G[k][k'] = sum over the B*H*W output pixels of patch_k * patch_k', k = (c, r, s): 171 products per pixel when done
// entry by entry (conv0_gram_kernel above).  But patch_k(h, w) = X[c][h + r - 1][w + s - 1], so over an output domain
// EXTENDED by one ring of pixels (h in [-1, H], w in [-1, W]; X zero outside the image) every tap sees every image
// pixel exactly once and the sum is a plain autocorrelation of the zero-padded input,
//     A[c][c'](r' - r, s' - s) = sum_{u,v} X[c][u][v] * X[c'][u + r' - r][v + s' - s],
// of which there are 13 (Cin = 1) or 51 (Cin = 2) distinct ones (A[c][c'](d) = A[c'][c](-d)).  What the ring adds is
// subtracted again: in its top row only the taps r = 2 are non-zero (image row 0), in its bottom row the taps r = 0, in
// its left / right columns the taps s = 2 / s = 0 -- four 1-D autocorrelations of the image's border rows / columns --
// and the four corners, taken away twice, come back once:
//     M G[(c,r,s)][(c',r',s')] = A - [r=r'=2] R_0(s'-s) - [r=r'=0] R_{H-1}(s'-s) - [s=s'=2] C_0(r'-r) - [s=s'=0] C_{W-1}(r'-r)
//                                  + [both taps at the same corner] X[c][corner] X[c'][corner]
// (same for the plain sums in column K).  51 + 2 products per pixel instead of 342, partials per block / per image,
// one warp per entry folds them in a fixed order in double precision: same `gram` as before to rounding.
template <int CIN>
struct AcDims {
    static constexpr int NA = CIN == 1 ? 13 : 51;        // distinct autocorrelation sums
    static constexpr int NACC = NA + CIN;                // + the plain sums S_c
    static constexpr int SEG = CIN * CIN * 5;            // one border row / column: [c][c'][d + 2]
    static constexpr int B1 = 4 * SEG;                   // border row / column plain sums [seg][c]
    static constexpr int B2 = B1 + 4 * CIN;              // corner values [corner][c]
    static constexpr int B3 = B2 + 4 * CIN;              // corner products [corner][c][c']
    static constexpr int NE = B3 + 4 * CIN * CIN;        // per-image edge partial
};
// index of A[c][c2](dr, ds) in a main partial (dr, ds in [-2, 2])
template <int CIN>
__host__ __device__ inline int ac_index(int c, int c2, int dr, int ds) {
    if (c == c2) {
        if (dr < 0 || (dr == 0 && ds < 0)) { dr = -dr; ds = -ds; }
        const int i = dr == 0 ? ds : 3 + (dr - 1) * 5 + (ds + 2);
        return c * 13 + i;
    }
    // c != c2 (CIN == 2): A01(dr >= 0) at 26 + dr*5 + ds+2 ; A10(dr > 0) = sum X1[u][v] X0[u+dr][v+ds] at 41 + (dr-1)*5 + ds+2
    if (c == 0) return dr >= 0 ? 26 + dr * 5 + (ds + 2) : 41 + (-dr - 1) * 5 + (-ds + 2);
    return dr > 0 ? 41 + (dr - 1) * 5 + (ds + 2) : 26 + (-dr) * 5 + (-ds + 2);
}

template <int CIN>
__device__ __forceinline__ void ac_edge_block(const float* __restrict__ x, int H, int W, int b, float* __restrict__ partE) {
    using D = AcDims<CIN>;
    constexpr int NW = D::SEG + CIN;                      // per-warp accumulators: the segment's products + plain sums
    __shared__ float sE[8][NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, seg = warp >> 1, half = warp & 1;
    const bool is_row = seg < 2;
    const int fixed = (seg == 0 || seg == 2) ? 0 : (is_row ? H - 1 : W - 1), len = is_row ? W : H;
    const float* img = x + (long)b * CIN * H * W;
    auto at = [&](int c, int t) -> float {                // element t of the border row / column of channel c, 0 outside
        if (t < 0 || t >= len) return 0.0f;
        return is_row ? __ldg(img + ((long)c * H + fixed) * W + t) : __ldg(img + ((long)c * H + t) * W + fixed);
    };
    float acc[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = 0.0f;
    for (int t = half * 32 + lane; t < len; t += 64) {
        float nb[CIN][5], own[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
#pragma unroll
            for (int d = 0; d < 5; ++d) nb[c][d] = at(c, t + d - 2);
            own[c] = nb[c][2];
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
#pragma unroll
            for (int c2 = 0; c2 < CIN; ++c2)
#pragma unroll
                for (int d = 0; d < 5; ++d) acc[(c * CIN + c2) * 5 + d] = fmaf(own[c], nb[c2][d], acc[(c * CIN + c2) * 5 + d]);
            acc[D::SEG + c] += own[c];
        }
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sE[warp][i] = v;
    }
    __syncthreads();
    float* out = partE + (long)b * D::NE;
    for (int i = threadIdx.x; i < 4 * NW; i += blockDim.x) {
        const int sg = i / NW, e = i - sg * NW;
        const float v = sE[2 * sg][e] + sE[2 * sg + 1][e];
        if (e < D::SEG) out[sg * D::SEG + e] = v;
        else out[D::B1 + sg * CIN + (e - D::SEG)] = v;
    }
    if (threadIdx.x < 4 * CIN) {                           // corner values, then their products
        const int corner = threadIdx.x / CIN, c = threadIdx.x - corner * CIN;
        const int u = (corner & 2) ? H - 1 : 0, v = (corner & 1) ? W - 1 : 0;
        const float xc = __ldg(img + ((long)c * H + u) * W + v);
        out[D::B2 + corner * CIN + c] = xc;
#pragma unroll
        for (int c2 = 0; c2 < CIN; ++c2) out[D::B3 + (corner * CIN + c) * CIN + c2] = xc * __ldg(img + ((long)c2 * H + u) * W + v);
    }
}

// blocks [0, n_img): the border rows / columns / corners of image b (first, so that they run beside the main blocks);
// blocks [n_img, n_img + n_main): persistent over row groups, 51 + 2 sums per thread.  The tile of the NEXT group is
// fetched into registers before the current one is evaluated (every thread owns the same <= kAcStage tile slots in every
// group, decoded once), so the global-load latency hides behind the products.
constexpr int kAcStage = 4;
template <int CIN>
__global__ void __launch_bounds__(256, 3)
conv0_ac_kernel(const float* __restrict__ x, int H, int W, unsigned wmagic, int groups_per_img, int n_groups, int n_img,
                float* __restrict__ partA, float* __restrict__ partE) {
    pdl_wait();
    using D = AcDims<CIN>;
    if ((int)blockIdx.x < n_img) { ac_edge_block<CIN>(x, H, W, (int)blockIdx.x, partE); return; }
    const int blk = (int)blockIdx.x - n_img, n_main = (int)gridDim.x - n_img;
    extern __shared__ float xs[];                          // [CIN][kC0Rows + 2][W + 4]: columns -2 .. W+1, zero outside
    __shared__ float red[8][D::NACC];
    const int Wp = W + 4, plane = (kC0Rows + 2) * Wp, n_slots = CIN * plane;
    // this thread's tile slots: (channel, tile row, image column) or "always zero"
    int s_off[kAcStage], s_rr[kAcStage];                   // offset inside the image of (ci, row 0, ww); -1: zero slot
#pragma unroll
    for (int j = 0; j < kAcStage; ++j) {
        const int i = threadIdx.x + j * 256;
        s_off[j] = -1; s_rr[j] = 0;
        if (i < n_slots) {
            const int ci = i / plane, rem = i - ci * plane, rr = rem / Wp, ww = rem - rr * Wp - 2;
            s_rr[j] = rr;
            if (ww >= 0 && ww < W) s_off[j] = ci * H * W + ww;
        }
    }
    auto fetch = [&](int b, int h0, float (&v)[kAcStage]) {
        const float* img = x + (long)b * CIN * H * W;
#pragma unroll
        for (int j = 0; j < kAcStage; ++j) {
            const int hh = h0 + s_rr[j];
            v[j] = (s_off[j] >= 0 && hh < H) ? __ldg(img + s_off[j] + hh * W) : 0.0f;
        }
    };
    float acc[D::NACC];
#pragma unroll
    for (int i = 0; i < D::NACC; ++i) acc[i] = 0.0f;
    GroupWalk gw(blk, n_main, groups_per_img);
    float cur[kAcStage], nxt[kAcStage];
    if (blk < n_groups) fetch(gw.b, gw.h0(), cur);
    for (int grp = blk; grp < n_groups; grp += n_main) {
        const int h0 = gw.h0();
#pragma unroll
        for (int j = 0; j < kAcStage; ++j)
            if (threadIdx.x + j * 256 < n_slots) xs[threadIdx.x + j * 256] = cur[j];
        gw.next();
        if (grp + n_main < n_groups) fetch(gw.b, gw.h0(), nxt);
        __syncthreads();
        const int npix = min(kC0Rows, H - h0) * W;
        for (int p = threadIdx.x; p < npix; p += blockDim.x) {
            const int row = (int)__umulhi((unsigned)p, wmagic), col = p - row * W;
            const float* b0 = xs + row * Wp + col + 2;     // channel 0 at (u, v); neighbours at [dr * Wp + ds]
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float* bc = b0 + c * plane;
                const float xc = bc[0];
                acc[D::NA + c] += xc;
                // A[c][c]: (0, 0..2), (1..2, -2..2)
#pragma unroll
                for (int ds = 0; ds <= 2; ++ds) acc[c * 13 + ds] = fmaf(xc, bc[ds], acc[c * 13 + ds]);
#pragma unroll
                for (int dr = 1; dr <= 2; ++dr)
#pragma unroll
                    for (int ds = -2; ds <= 2; ++ds)
                        acc[c * 13 + 3 + (dr - 1) * 5 + ds + 2] = fmaf(xc, bc[dr * Wp + ds], acc[c * 13 + 3 + (dr - 1) * 5 + ds + 2]);
                if (CIN == 2) {
                    const float* bo = b0 + (1 - c) * plane;   // the OTHER channel
                    if (c == 0) {                          // A01(dr = 0..2, ds): X0[u][v] * X1[u + dr][v + ds]
#pragma unroll
                        for (int dr = 0; dr <= 2; ++dr)
#pragma unroll
                            for (int ds = -2; ds <= 2; ++ds)
                                acc[26 + dr * 5 + ds + 2] = fmaf(xc, bo[dr * Wp + ds], acc[26 + dr * 5 + ds + 2]);
                    } else {                               // A10(dr = 1..2, ds): X1[u][v] * X0[u + dr][v + ds]
#pragma unroll
                        for (int dr = 1; dr <= 2; ++dr)
#pragma unroll
                            for (int ds = -2; ds <= 2; ++ds)
                                acc[41 + (dr - 1) * 5 + ds + 2] = fmaf(xc, bo[dr * Wp + ds], acc[41 + (dr - 1) * 5 + ds + 2]);
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kAcStage; ++j) cur[j] = nxt[j];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < D::NACC; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < D::NACC) {
        float t = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
        partA[(long)blk * D::NACC + threadIdx.x] = t;
    }
}

// gram[k][k'] (k' <= K, doubles, divided by the pixel count) from the autocorrelation / border partials: one warp per
// entry, every term folded over its blocks / images in a fixed order (lanes take every 32nd, fixed shuffle tree)
template <int CIN>
__global__ void __launch_bounds__(256)
conv0_ac_gram_kernel(const float* __restrict__ partA, int n_main, const float* __restrict__ partE, int n_img, double inv_n,
                     double* __restrict__ gram) {
    pdl_wait();
    using D = AcDims<CIN>;
    constexpr int K = CIN * 9, NEnt = K * (K + 1);
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= NEnt) return;
    const int k = e / (K + 1), k2 = e - k * (K + 1);
    const int c = k / 9, r = (k % 9) / 3, s = k % 3;
    auto fold = [&](const float* part, int stride, int n, int idx) -> double {
        const int mine = n > lane ? (n - lane + 31) / 32 : 0;
        double a = ordered_sum<8, double>(part + (long)lane * stride + idx, 32L * stride, mine);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        return a;
    };
    auto foldA = [&](int idx) { return fold(partA, D::NACC, n_main, idx); };
    auto foldE = [&](int idx) { return fold(partE, D::NE, n_img, idx); };
    // corner index of a tap (r, s): (2,2) -> top-left 0, (2,0) -> top-right 1, (0,2) -> bottom-left 2, (0,0) -> bottom-right 3
    auto corner_of = [](int rr, int ss) { return (rr == 1 || ss == 1) ? -1 : ((rr == 0 ? 2 : 0) + (ss == 0 ? 1 : 0)); };
    double val;
    if (k2 == K) {                                         // plain sum of patch_k
        val = foldA(D::NA + c);
        if (r == 2) val -= foldE(D::B1 + 0 * CIN + c);
        if (r == 0) val -= foldE(D::B1 + 1 * CIN + c);
        if (s == 2) val -= foldE(D::B1 + 2 * CIN + c);
        if (s == 0) val -= foldE(D::B1 + 3 * CIN + c);
        const int cr = corner_of(r, s);
        if (cr >= 0) val += foldE(D::B2 + cr * CIN + c);
    } else {
        const int c2 = k2 / 9, r2 = (k2 % 9) / 3, s2 = k2 % 3, dr = r2 - r, ds = s2 - s;
        val = foldA(ac_index<CIN>(c, c2, dr, ds));
        if (r == 2 && r2 == 2) val -= foldE(0 * D::SEG + (c * CIN + c2) * 5 + ds + 2);
        if (r == 0 && r2 == 0) val -= foldE(1 * D::SEG + (c * CIN + c2) * 5 + ds + 2);
        if (s == 2 && s2 == 2) val -= foldE(2 * D::SEG + (c * CIN + c2) * 5 + dr + 2);
        if (s == 0 && s2 == 0) val -= foldE(3 * D::SEG + (c * CIN + c2) * 5 + dr + 2);
        const int cr = corner_of(r, s);
        if (cr >= 0 && cr == corner_of(r2, s2)) val += foldE(D::B3 + (cr * CIN + c) * CIN + c2);
    }
    if (lane == 0) gram[e] = val * inv_n;
}

// gram[k][k'] (k' <= K, doubles, divided by the pixel count): one warp per entry, fixed order
__global__ void conv0_gram_reduce_kernel(const float* __restrict__ part, int nblk, int cin, double inv_n,
                                         double* __restrict__ gram) {
    pdl_wait();
    const int K = cin * 9, NE = K * (K + 1);
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= NE) return;
    const int mine = nblk > lane ? (nblk - lane + 31) / 32 : 0;
    double a = ordered_sum<8, double>(part + (long)lane * NE + e, 32L * NE, mine);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) gram[e] = a * inv_n;
}

// BatchNorm statistics of conv 0 from the patch moments; same outputs as bn_finalize_train_kernel.  One warp per
// channel: lane k owns row k of w^T (G - m m^T) w, fixed shuffle tree.
__global__ void __launch_bounds__(256)
conv0_bn_finalize_kernel(const double* __restrict__ gram, int cin, int C, long n, const float* __restrict__ w,
                         const float* __restrict__ bias, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running,
                         float* __restrict__ stat) {
    pdl_wait();
    __shared__ double gsh[18 * 19];
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, k = threadIdx.x & 31, K = cin * 9;
    for (int i = threadIdx.x; i < K * (K + 1); i += blockDim.x) gsh[i] = gram[i];
    __syncthreads();
    gram = gsh;
    if (c >= C) return;
    const float* wc = w + (long)c * K;
    double mpart = 0.0, vpart = 0.0;
    if (k < K) {
        const double mk = gram[k * (K + 1) + K], wk = (double)wc[k];
        double row = 0.0;
        for (int k2 = 0; k2 < K; ++k2) row += (double)wc[k2] * (gram[k * (K + 1) + k2] - mk * gram[k2 * (K + 1) + K]);
        mpart = wk * mk;
        vpart = wk * row;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mpart += __shfl_xor_sync(0xffffffffu, mpart, o);
        vpart += __shfl_xor_sync(0xffffffffu, vpart, o);
    }
    if (k != 0) return;
    const double mean = (double)bias[c] + mpart;
    double var = vpart;
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

// conv + bias -> BN -> ReLU -> max-pool(1,P) -> dropout in registers.  Same accumulation order as
// conv0_fwd_stats_kernel (y is bit-identical to what that kernel would have stored).  lane <-> 4 output channels
// (weights in registers as channel pairs for fma.f32x2), warp <-> one image row at a time.  Every warp stages ITS
// OWN three input rows (per input channel) with cp.async, one row ahead, and never meets a block-wide barrier: the
// warps of an SM drift apart, so the FMA-heavy window bodies of some overlap the max / dropout / store tails of others.
template <int CIN, int P>
__global__ void __launch_bounds__(256, 2)
conv0_lean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ stat, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                      __nv_bfloat16* __restrict__ out_lo, unsigned* __restrict__ argw, PoolGeom g, int n_rows) {
    pdl_wait();
    extern __shared__ float xs_all[];                     // [8 warps][2][CIN][3][W+2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = blockIdx.y * 32 + lane, c = c4 * 4, C4 = g.C >> 2, Wp = g.W + 2, xsz = CIN * 3 * Wp;
    float* xw = xs_all + warp * 2 * xsz;
    float2 wr[2][CIN * 9], bs[2], sc[2], sh[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        bs[q] = make_float2(__ldg(bias + c + 2 * q), __ldg(bias + c + 2 * q + 1));
        sc[q] = make_float2(__ldg(stat + 2 * g.C + c + 2 * q), __ldg(stat + 2 * g.C + c + 2 * q + 1));
        sh[q] = make_float2(__ldg(stat + 3 * g.C + c + 2 * q), __ldg(stat + 3 * g.C + c + 2 * q + 1));
#pragma unroll
        for (int k = 0; k < CIN * 9; ++k)
            wr[q][k] = make_float2(__ldg(w + (long)(c + 2 * q) * CIN * 9 + k), __ldg(w + (long)(c + 2 * q + 1) * CIN * 9 + k));
    }
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    // rows h-1, h, h+1 of every input channel of image b, zero-filled outside the image
    auto stage_row = [&](float* dst, int b, int h) {
#pragma unroll
        for (int rowid = 0; rowid < CIN * 3; ++rowid) {
            const int ci = rowid / 3, hh = h - 1 + (rowid - ci * 3);
            const bool rowok = hh >= 0 && hh < g.H;
            const float* src = x + (((long)b * CIN + ci) * g.H + (rowok ? hh : 0)) * g.W;
            for (int cc = lane; cc < Wp; cc += 32) {
                const bool ok = rowok && cc >= 1 && cc <= g.W;
                cpa4_zfill(dst + rowid * Wp + cc, src + (ok ? cc - 1 : 0), ok);
            }
        }
    };
    const int stride = gridDim.x * kC0Rows;
    int row = blockIdx.x * kC0Rows + warp, buf = 0;
    int b = row / g.H, h = row - b * g.H;
    const int step_b = stride / g.H, step_h = stride - step_b * g.H;
    if (row < n_rows) stage_row(xw, b, h);
    cpa_commit();
    for (; row < n_rows; row += stride, buf ^= 1) {
        int bn = b + step_b, hn = h + step_h;
        if (hn >= g.H) { hn -= g.H; ++bn; }
        if (row + stride < n_rows) stage_row(xw + (buf ^ 1) * xsz, bn, hn);
        cpa_commit();
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const float* xs = xw + buf * xsz;
        for (int wo = 0; wo < g.Wo; ++wo) {
            float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            unsigned arg[4] = {0u, 0u, 0u, 0u};
            auto consider = [&](int j, const float2 a01, const float2 a23) {
                const float2 z01 = __ffma2_rn(a01, sc[0], sh[0]), z23 = __ffma2_rn(a23, sc[1], sh[1]);
                const float z[4] = {z01.x, z01.y, z23.x, z23.y};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (z[q] > best[q]) { best[q] = z[q]; arg[q] = (unsigned)j; }      // first maximum wins
            };
#pragma unroll
            for (int j = 0; j + 1 < P; j += 2) {
                const int ww = wo * P + j;
                float2 acc[2][2] = {{bs[0], bs[1]}, {bs[0], bs[1]}};
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float* xr = xs + (ci * 3 + r) * Wp + ww;
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
                consider(j, acc[0][0], acc[0][1]);
                consider(j + 1, acc[1][0], acc[1][1]);
            }
            if (P & 1) {
                const int ww = wo * P + P - 1;
                float2 acc[2] = {bs[0], bs[1]};
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const float xv = xs[(ci * 3 + r) * Wp + ww + t];
                            const float2 x2 = make_float2(xv, xv);
                            acc[0] = __ffma2_rn(wr[0][ci * 9 + r * 3 + t], x2, acc[0]);
                            acc[1] = __ffma2_rn(wr[1][ci * 9 + r * 3 + t], x2, acc[1]);
                        }
                consider(P - 1, acc[0], acc[1]);
            }
            const long pix = ((long)b * g.H + h) * g.Wo + wo;
            const long i = pix * C4 + c4;                      // element numbering of the pool kernels (dropout)
            Keep4 kp;
#pragma unroll
            for (int q = 0; q < 4; ++q) kp.k[q] = true;
            if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
            float m[4];
            unsigned word = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool alive = best[q] > 0.0f && kp.k[q];
                m[q] = alive ? best[q] * keep_scale : 0.0f;
                word |= (arg[q] | (alive ? 0u : 0x80u)) << (8 * q);
            }
            const float4 m4 = make_float4(m[0], m[1], m[2], m[3]);
            if (argw) argw[i] = word;
            if (out_hi) store_planes4(out_hi, out_lo, pix, c4, C4, m4);
            if (out) {
                float* dst = out + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC;
                if (g.oC == 1) {
                    *reinterpret_cast<float4*>(dst) = m4;
                } else {
                    dst[0] = m[0]; dst[g.oC] = m[1]; dst[2 * g.oC] = m[2]; dst[3 * g.oC] = m[3];
                }
            }
        }
        __syncwarp();                                          // every lane is done with xs[buf] before it is refilled
        b = bn; h = hn;
    }
}

// S[k][c] = sum over windows of dz * patch_k(winner), S[K][c] = sum dz; part layout [nblk][K+1][C].
// dA and the winner bytes of the NEXT tile (8 rows x kLeanWC windows x 128 channels) travel global -> shared with
// 16-byte cp.async while the current tile is consumed, so no warp ever waits on HBM; channel pairs accumulate with
// packed fma.f32x2 (each half has its own winner, i.e. its own patch value).  Needs the channels-last block output
// (g.oC == 1: block 0 is not the last conv block).
constexpr int kLeanWC = 8;
inline size_t conv0_lean_bwd_smem(int cin, int W) {
    const size_t xpad = (2 * (size_t)cin * (kC0Rows + 2) * (W + 2) + 3) & ~(size_t)3;
    return xpad * 4 + 2 * (size_t)kC0Rows * kLeanWC * (128 + 32) * 4;
}
template <int CIN, int P>
__global__ void __launch_bounds__(256)
conv0_lean_bwd_kernel(const float* __restrict__ x, const unsigned* __restrict__ argw, const float* __restrict__ dA,
                      PoolGeom g, int groups_per_img, int n_groups, float* __restrict__ part) {
    pdl_wait();
    constexpr int K = CIN * 9;
    extern __shared__ __align__(16) float lean_smem[];
    __shared__ float red[kC0Rows][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C4 = g.C >> 2, Wp = g.W + 2, xsz = CIN * (kC0Rows + 2) * Wp;
    const int xpad = (2 * xsz + 3) & ~3;                                         // keep the tiles 16 B aligned
    float* xs_all = lean_smem;                                                   // 2 x [CIN][10][W+2]
    float* da_all = lean_smem + xpad;                                            // 2 x [8][WC][128]
    unsigned* ar_all = reinterpret_cast<unsigned*>(da_all + 2 * kC0Rows * kLeanWC * 128);   // 2 x [8][WC][32]
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    const int n_chunks = (g.Wo + kLeanWC - 1) / kLeanWC;
    float2 acc[2][K + 1];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k <= K; ++k) acc[q][k] = make_float2(0.0f, 0.0f);

    // one warp per image row of the tile: 512 B of dA (all lanes) + 128 B of winner words (lanes 0..7) per window
    auto stage_tile = [&](int tb, int b, int h0, int chunk) {
        const int h = h0 + warp;
        if (h >= g.H) return;
        const int w0 = chunk * kLeanWC, nw = min(kLeanWC, g.Wo - w0);
        const float* src = dA + (long)b * g.oB + (long)h * g.oH + (long)w0 * g.oW + blockIdx.y * 128 + lane * 4;
        const unsigned* asrc = argw + (((long)b * g.H + h) * g.Wo + w0) * C4 + blockIdx.y * 32 + lane * 4;
        float* dd = da_all + ((tb * kC0Rows + warp) * kLeanWC) * 128 + lane * 4;
        unsigned* ad = ar_all + ((tb * kC0Rows + warp) * kLeanWC) * 32 + lane * 4;
        for (int wl = 0; wl < nw; ++wl) {
            cpa16(dd + wl * 128, src + (long)wl * g.oW);
            if (lane < 8) cpa16(ad + wl * 32, asrc + (long)wl * C4);
        }
    };

    int grp = blockIdx.x, xb = 0, tb = 0;
    GroupWalk gw(grp, gridDim.x, groups_per_img), gn = gw;
    if (grp < n_groups) {
        stage_x_rows_async<CIN>(x, xs_all, gw.b, gw.h0(), g.H, g.W);
        stage_tile(0, gw.b, gw.h0(), 0);
    }
    cpa_commit();
    cpa_wait_all();
    __syncthreads();
    for (; grp < n_groups; grp += gridDim.x, xb ^= 1) {
        const float* xs = xs_all + xb * xsz;
        const int h0 = gw.h0();
        gn.next();
        const bool more = grp + (int)gridDim.x < n_groups;
        for (int chunk = 0; chunk < n_chunks; ++chunk, tb ^= 1) {
            if (chunk + 1 < n_chunks) {
                stage_tile(tb ^ 1, gw.b, h0, chunk + 1);
            } else if (more) {
                stage_x_rows_async<CIN>(x, xs_all + (xb ^ 1) * xsz, gn.b, gn.h0(), g.H, g.W);
                stage_tile(tb ^ 1, gn.b, gn.h0(), 0);
            }
            cpa_commit();
            if (h0 + warp < g.H) {
                const int w0 = chunk * kLeanWC, nw = min(kLeanWC, g.Wo - w0);
                const float* xrow = xs + warp * Wp + w0 * P;
                const float* dd = da_all + ((tb * kC0Rows + warp) * kLeanWC) * 128 + lane * 4;
                const unsigned* ad = ar_all + ((tb * kC0Rows + warp) * kLeanWC) * 32 + lane;
                for (int wl = 0; wl < nw; ++wl) {
                    const float4 g4 = *reinterpret_cast<const float4*>(dd + wl * 128);
                    const unsigned word = ad[wl * 32];
                    const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        const unsigned b0 = (word >> (16 * pr)) & 0xFFu, b1 = (word >> (16 * pr + 8)) & 0xFFu;
                        const float2 dz = make_float2((b0 & 0x80u) ? 0.0f : gq[2 * pr] * keep_scale,
                                                      (b1 & 0x80u) ? 0.0f : gq[2 * pr + 1] * keep_scale);
                        const float* base0 = xrow + wl * P + (int)(b0 & 0x7Fu);
                        const float* base1 = xrow + wl * P + (int)(b1 & 0x7Fu);
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int t = 0; t < 3; ++t) {
                                    const int off = (ci * (kC0Rows + 2) + r) * Wp + t;
                                    acc[pr][ci * 9 + r * 3 + t] = __ffma2_rn(dz, make_float2(base0[off], base1[off]),
                                                                             acc[pr][ci * 9 + r * 3 + t]);
                                }
                        acc[pr][K] = __fadd2_rn(acc[pr][K], dz);
                    }
                }
            }
            cpa_wait_all();
            __syncthreads();
        }
        gw = gn;
    }
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        red[warp][lane * 4 + 0] = acc[0][k].x; red[warp][lane * 4 + 1] = acc[0][k].y;
        red[warp][lane * 4 + 2] = acc[1][k].x; red[warp][lane * 4 + 3] = acc[1][k].y;
        __syncthreads();
        if (threadIdx.x < 128) {
            float t = 0.0f;
#pragma unroll
            for (int r = 0; r < kC0Rows; ++r) t += red[r][threadIdx.x];
            part[((long)blockIdx.x * (K + 1) + k) * g.C + blockIdx.y * 128 + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// S[k][c] = sum_blk part[blk][k][c] in double: block <-> (k, 128-channel slice), thread <-> (channel, 1 of 8 interleaved
// block ranges); every load is a coalesced 512 B row, the eight range sums are added in a fixed order
__global__ void __launch_bounds__(1024)
conv0_lean_bwd_colsum_kernel(const float* __restrict__ part, int nblk, int K1, int C, double* __restrict__ S) {
    pdl_wait();
    __shared__ double red[8][128];
    const int k = blockIdx.x, c = blockIdx.y * 128 + (threadIdx.x & 127), grp = threadIdx.x >> 7;
    const int mine = nblk > grp ? (nblk - grp + 7) / 8 : 0;
    const double a = ordered_sum<8, double>(part + ((long)grp * K1 + k) * C + c, 8L * K1 * C, mine);
    red[grp][threadIdx.x & 127] = a;
    __syncthreads();
    if (grp == 0) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
        S[(long)k * C + c] = t;
    }
}

// d(beta), d(gamma), dW of conv 0 (and its zero bias gradient) from S, the patch moments and the forward statistics;
// one warp per channel, lane k <-> dW[c][k]
__global__ void __launch_bounds__(256)
conv0_lean_bwd_finalize_kernel(const double* __restrict__ S, int cin, int C, const double* __restrict__ gram,
                               const float* __restrict__ w, const float* __restrict__ bias,
                               const float* __restrict__ gamma, const float* __restrict__ stat,
                               float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dgamma,
                               float* __restrict__ dbeta) {
    pdl_wait();
    __shared__ double gsh[18 * 19];
    const int K = cin * 9;
    for (int i = threadIdx.x; i < K * (K + 1); i += blockDim.x) gsh[i] = gram[i];
    __syncthreads();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, k = threadIdx.x & 31;
    if (c >= C || k >= K) return;
    const float* wc = w + (long)c * K;
    const double mean = (double)stat[c], invstd = (double)stat[C + c], bc = (double)bias[c];
    const double sum_dz = S[(long)K * C + c];
    double sum_dz_y = bc * sum_dz, wg = 0.0;
    for (int k2 = 0; k2 < K; ++k2) {
        sum_dz_y += (double)wc[k2] * S[(long)k2 * C + c];
        wg += (double)wc[k2] * gsh[k2 * (K + 1) + k];                  // G is symmetric
    }
    const double dg = invstd * (sum_dz_y - mean * sum_dz);             // sum dz * xhat
    const double mk = gsh[k * (K + 1) + K];
    const double tk = invstd * (wg + (bc - mean) * mk);                // mean over pixels of xhat * patch_k
    dw[(long)c * K + k] = (float)((double)gamma[c] * invstd * (S[(long)k * C + c] - sum_dz * mk - dg * tk));
    if (k == 0) {
        dgamma[c] = (float)dg;
        dbeta[c] = (float)sum_dz;
        db[c] = 0.0f;
    }
}


// ----------------------------------------------------------------------------- lean block 0 forward on tcgen05
// Same contract as conv0_lean_fwd_kernel (pooled output + winner bytes, statistics known up front), with the
// contraction on the tensor cores: a tile is NW = 128/P pooling windows = TP = NW*P pixels (rows of the MMA),
//   1. two threads (warps 0-3: hi plane, warps 4-7: lo plane) build one im2col row (K = 9*Cin <= 18 values, zero-padded to 16-wide k-steps) as bf16 hi / lo
//      planes straight into the 128-byte-swizzled K-major layout tcgen05 reads (what TMA would have produced),
//   2. one thread issues 3 MMAs (hi*hi + hi*lo + lo*hi, the fp32-grade split of tc_conv.cu) per k-step into a
//      128 x 128 fp32 TMEM accumulator,
//   3. the 8 warps move the accumulator TMEM -> registers -> an XOR-swizzled [pixel][channel] shared tile (it aliases
//      the A planes, which the MMAs are done with),
//   4. warp <-> window, lane <-> 4 channels: bias + BN + max / argmax over the window's P rows, ReLU, dropout, stores.
// Two CTAs per SM (96 KB of shared memory, 128 TMEM columns each) overlap one CTA's epilogue with the other's build.
namespace c0tc {
constexpr int kPlane = 128 * 128;                               // bytes: 128 rows x 128 B (64 bf16 K slots, <= 32 used)
constexpr int kSmem = 1024 + 2 * kPlane + 128 * 128 * 4 + 18 * 128 * 4 + 64;   // align + B planes + tile (aliases the A planes) + patch staging + barrier
__device__ __forceinline__ uint32_t sw128(int r, int c) {       // byte offset of 16-byte chunk c of row r
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
template <int K>
__device__ __forceinline__ void put_row(unsigned char* plane, int r, const float (&v)[K], bool lo) {
    constexpr int CH = ((K + 15) / 16) * 2;                     // chunks of 8 bf16 covering the k-steps in use
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = c * 8 + 2 * e;
            const float a = k0 < K ? v[k0 < K ? k0 : 0] : 0.0f, b = k0 + 1 < K ? v[k0 + 1 < K ? k0 + 1 : 0] : 0.0f;
            __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
            if (lo) {
                ha = __float2bfloat16_rn(a - __bfloat162float(ha));
                hb = __float2bfloat16_rn(b - __bfloat162float(hb));
            }
            w[e] = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
        }
        *reinterpret_cast<uint4*>(plane + sw128(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
}  // namespace c0tc

template <int CIN, int P>
__global__ void __launch_bounds__(256, 2)
conv0_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    const float* __restrict__ stat, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                    __nv_bfloat16* __restrict__ out_lo, unsigned* __restrict__ argw, PoolGeom g, unsigned n_windows,
                    int n_tiles) {
    pdl_wait();
    using namespace umma;
    constexpr int K = CIN * 9, KSTEPS = (K + 15) / 16, NW = 128 / P, TP = NW * P;
    constexpr int KH = (K + 1) / 2;                               // patch entries fetched by each thread of a row pair
    extern __shared__ unsigned char c0tc_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c0tc_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *b_hi = smem, *b_lo = smem + c0tc::kPlane;
    unsigned char *a_hi = smem + 2 * c0tc::kPlane, *a_lo = a_hi + c0tc::kPlane;
    float* ys = reinterpret_cast<float*>(smem + 2 * c0tc::kPlane);                   // [128][128], aliases a_hi / a_lo
    float* xst = reinterpret_cast<float*>(smem + 2 * c0tc::kPlane + 128 * 128 * 4);  // [K][128] patch staging (cp.async)
    uint64_t* bar = reinterpret_cast<uint64_t*>(xst + 18 * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    {   // B operand: row n = output channel, the same K ordering as the im2col rows (w[c][ci][r][t])
        const int n = tid & 127;
        float v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldg(w + (long)(blockIdx.y * 128 + n) * K + k);
        c0tc::put_row<K>((tid >> 7) ? b_lo : b_hi, n, v, (tid >> 7) != 0);
    }
    // patch values of a tile's rows: global -> shared with zero-filling 4-byte cp.async, one tile ahead of their use;
    // the two threads of a row (tid, tid + 128) fetch half of its entries each
    const unsigned uWo = (unsigned)g.Wo, uH = (unsigned)g.H;
    auto fetch_tile = [&](int tile) {
        const int r = tid & 127;
        const unsigned q = (unsigned)tile * NW + (unsigned)(r / P);
        const int j = r % P;
        const bool rowok = r < TP && q < n_windows;
        const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
        const int wc = (int)wo * P + j;
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
            const int k = (tid >> 7) * KH + kk;
            if (k < K) {
                const int ci = k / 9, rr = (k - ci * 9) / 3, t = k - ci * 9 - rr * 3;
                const int hh = (int)h + rr - 1, ww = wc + t - 1;
                const bool ok = rowok && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
                const float* src = x + (((long)b * CIN + ci) * g.H + (ok ? hh : 0)) * g.W + (ok ? ww : 0);
                cpa4_zfill(xst + k * 128 + r, ok ? src : x, ok);
            }
        }
    };
    if ((int)blockIdx.x < n_tiles) fetch_tile(blockIdx.x);
    cpa_commit();
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int c4 = blockIdx.y * 32 + lane, c = c4 * 4, C4 = g.C >> 2;
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias + c));
    const float4 sc = __ldg(reinterpret_cast<const float4*>(stat + 2 * g.C + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(stat + 3 * g.C + c));
    // z = sc * (acc + bias) + sh: the conv bias is folded into the shift
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
    const float shv[4] = {fmaf(bs.x, sc.x, sh.x), fmaf(bs.y, sc.y, sh.y), fmaf(bs.z, sc.z, sh.z), fmaf(bs.w, sc.w, sh.w)};
    const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- 1. im2col rows of this tile (thread <-> row; warps 0-3: hi plane, warps 4-7: lo plane)
        cpa_wait_all();
        __syncthreads();                                       // staged patches visible; previous tile's windows are done
        {
            const int r = tid & 127;
            float v[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = xst[k * 128 + r];
            c0tc::put_row<K>((tid >> 7) ? a_lo : a_hi, r, v, (tid >> 7) != 0);
        }
        fence_proxy_async();
        __syncthreads();
        // ---- 2. MMAs
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = idesc_bf16(128, 128, 0, 0);
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh2 = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t dah = smem_desc_sw128(ah + k * 32, 16, 1024), dal = smem_desc_sw128(al + k * 32, 16, 1024);
                const uint64_t dbh = smem_desc_sw128(bh2 + k * 32, 16, 1024), dbl = smem_desc_sw128(bl + k * 32, 16, 1024);
                mma_bf16(tmem, dah, dbh, idesc, k != 0);
                mma_bf16(tmem, dah, dbl, idesc, 1);
                mma_bf16(tmem, dal, dbh, idesc, 1);
            }
            mma_commit(bar);
        }
        if (tile + (int)gridDim.x < n_tiles) fetch_tile(tile + gridDim.x);     // lands while the MMAs and the epilogue run
        cpa_commit();
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- 3. accumulator -> shared [pixel][channel] tile (chunk index XOR row: conflict-free both ways)
        {
            const int qd = warp & 3, half = warp >> 2, r = qd * 32 + lane;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + half * 64 + cc * 32, v);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int chunk = half * 16 + cc * 8 + jj;
                    *reinterpret_cast<float4*>(ys + r * 128 + ((chunk ^ lane) << 2)) =
                        make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- 4. windows: warp <-> window, lane <-> 4 channels
        for (int wl = warp; wl < NW; wl += 8) {
            const unsigned q = (unsigned)tile * NW + (unsigned)wl;
            if (q >= n_windows) break;
            float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            unsigned arg[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const int row = wl * P + j;
                const float4 y4 = *reinterpret_cast<const float4*>(ys + row * 128 + ((lane ^ (row & 31)) << 2));
                const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const float z = fmaf(yv[qq], scv[qq], shv[qq]);
                    if (z > best[qq]) { best[qq] = z; arg[qq] = (unsigned)j; }      // first maximum wins
                }
            }
            const long i = (long)q * C4 + c4;                  // element numbering of the pool kernels (dropout)
            Keep4 kp;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) kp.k[qq] = true;
            if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
            float m[4];
            unsigned word = 0;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const bool alive = best[qq] > 0.0f && kp.k[qq];
                m[qq] = alive ? best[qq] * keep_scale : 0.0f;
                word |= (arg[qq] | (alive ? 0u : 0x80u)) << (8 * qq);
            }
            const float4 m4 = make_float4(m[0], m[1], m[2], m[3]);
            if (argw) argw[i] = word;
            if (out_hi) store_planes4(out_hi, out_lo, (long)q, c4, C4, m4);
            if (out) {
                const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
                float* dst = out + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)c * g.oC;
                if (g.oC == 1) {
                    *reinterpret_cast<float4*>(dst) = m4;
                } else {
                    dst[0] = m[0]; dst[g.oC] = m[1]; dst[2 * g.oC] = m[2]; dst[3 * g.oC] = m[3];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}


}  // namespace

bool conv0_lean_ok(int cin, int C, int pool, int n_conv) {
    return conv0_direct_ok(cin, C) && (pool == 5 || pool == 2) && n_conv > 1;
}

// BatchNorm batch statistics of conv 0's output from the patch moments of the input (training forward): fills `gram`
// ([K][K+1] doubles, kept for the backward pass), `stat` ([4][C]) and updates the running statistics
int conv0_lean_stats(const float* x, int cin, int C, int H, int W, int batch, const float* w, const float* bias,
                     const float* gamma, const float* beta, float eps, float momentum, float* running, float* stat,
                     double* gram, float* part, cudaStream_t st) {
    const int gpi = (H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch, K0 = 9 * cin;
    const long M = (long)batch * H * W;
    const unsigned wmagic = (unsigned)(((1ull << 32) + (unsigned)W - 1) / (unsigned)W);
    const int entries = K0 * (K0 + 1);
    const char* e_direct = std::getenv("SEDB200_GRAM_DIRECT");        // read at every call: the tests run both paths
    const bool direct = e_direct && std::atoi(e_direct) != 0;
    if (!direct && kC0Rows * W < 65536 && (size_t)cin * (kC0Rows + 2) * (W + 4) <= (size_t)kAcStage * 256) {
        // patch moments from the autocorrelations of the input (51 + 2 products per pixel instead of 342); one wave of
        // three blocks per SM: the per-image border blocks first, the rest walks the row groups
        const int n_main = std::min(n_groups, std::max(sm_count(), 3 * sm_count() - batch));
        const int nacc = cin == 1 ? AcDims<1>::NACC : AcDims<2>::NACC, ne = cin == 1 ? AcDims<1>::NE : AcDims<2>::NE;
        float* partA = part;
        float* partE = part + (size_t)n_main * nacc;
        const size_t asm_bytes = (size_t)cin * (kC0Rows + 2) * (W + 4) * 4;
        if (cin == 1) launch_k(conv0_ac_kernel<1>, n_main + batch, 256, asm_bytes, st, x, H, W, wmagic, gpi, n_groups, batch, partA, partE);
        else launch_k(conv0_ac_kernel<2>, n_main + batch, 256, asm_bytes, st, x, H, W, wmagic, gpi, n_groups, batch, partA, partE);
        SED_POST_LAUNCH();
        (void)ne;
        if (cin == 1) launch_k(conv0_ac_gram_kernel<1>, (entries * 32 + 255) / 256, 256, 0, st, partA, n_main, partE, batch, 1.0 / (double)M, gram);
        else launch_k(conv0_ac_gram_kernel<2>, (entries * 32 + 255) / 256, 256, 0, st, partA, n_main, partE, batch, 1.0 / (double)M, gram);
        SED_POST_LAUNCH();
    } else {
    const int gblk = std::min(n_groups, (cin == 1 ? 5 : 3) * sm_count());
    const size_t gsm = 2 * (size_t)cin * (kC0Rows + 2) * (W + 32) * 4;
    if (cin == 1) launch_k(conv0_gram_kernel<1>, gblk, GramDims<1>::THREADS, gsm, st, x, H, W, wmagic, gpi, n_groups, part);
    else launch_k(conv0_gram_kernel<2>, gblk, GramDims<2>::THREADS, gsm, st, x, H, W, wmagic, gpi, n_groups, part);
    SED_POST_LAUNCH();
    launch_k(conv0_gram_reduce_kernel, (entries * 32 + 255) / 256, 256, 0, st, part, gblk, cin, 1.0 / (double)M, gram);
    SED_POST_LAUNCH();
    }
    launch_k(conv0_bn_finalize_kernel, (C * 32 + 255) / 256, 256, 0, st, gram, cin, C, M, w, bias, gamma, beta, eps, momentum,
                                                                  running, stat);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// conv + BN (scale / shift in `stat`) + ReLU + max-pool + dropout in one kernel; writes the pooled output (fp32 `out`
// with the strides of `g`, and / or bf16 planes) and, when `argw` is given, the winner bytes for the backward pass
int conv0_lean_forward(const float* x, int cin, int batch, const float* w, const float* bias, const float* stat,
                       const PoolGeom& g, int tensor_cores, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                       unsigned* argw, cudaStream_t st) {
    const int gpi = (g.H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch;
    if (tensor_cores) {
        // contraction on tcgen05 (same 3-term split as conv 2 / 3); the fp32 kernel below serves tensor_cores = 0.
        // Window-row formulation (conv0_win.cu) unless SEDB200_CONV0_PIXEL_ROWS=1 asks for the pixel-row kernel below
        static const bool pixel_rows = [] { const char* e = std::getenv("SEDB200_CONV0_PIXEL_ROWS"); return e && e[0] == '1'; }();
        if (!pixel_rows && conv0_win_ok(cin, g.C, g.p, (long)batch * g.H * g.Wo))
            return conv0_win_forward(x, cin, batch, w, bias, stat, g, out, out_hi, out_lo, argw, st);
        const unsigned n_windows = (unsigned)((long)batch * g.H * g.Wo);
        const int NWt = 128 / g.p;
        const int n_tiles = (int)((n_windows + (unsigned)NWt - 1) / (unsigned)NWt);
        const dim3 tgrid(std::min(n_tiles, 2 * sm_count()), g.C / 128);
        const void* kfn = cin == 1 ? (g.p == 5 ? (const void*)conv0_tc_fwd_kernel<1, 5> : (const void*)conv0_tc_fwd_kernel<1, 2>)
                                   : (g.p == 5 ? (const void*)conv0_tc_fwd_kernel<2, 5> : (const void*)conv0_tc_fwd_kernel<2, 2>);
        const int rc = ensure_dyn_smem(kfn, c0tc::kSmem);
        if (rc) return rc;
        if (cin == 1 && g.p == 5) launch_k(conv0_tc_fwd_kernel<1, 5>, tgrid, 256, c0tc::kSmem, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_windows, n_tiles);
        else if (cin == 1) launch_k(conv0_tc_fwd_kernel<1, 2>, tgrid, 256, c0tc::kSmem, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_windows, n_tiles);
        else if (g.p == 5) launch_k(conv0_tc_fwd_kernel<2, 5>, tgrid, 256, c0tc::kSmem, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_windows, n_tiles);
        else launch_k(conv0_tc_fwd_kernel<2, 2>, tgrid, 256, c0tc::kSmem, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_windows, n_tiles);
        SED_POST_LAUNCH();
        return SEDB200_OK;
    }
    const dim3 grid(std::min(n_groups, 2 * sm_count()), g.C / 128);
    const size_t smw = (size_t)kC0Rows * 2 * cin * 3 * (g.W + 2) * 4;      // per-warp row buffers
    const int n_rows = batch * g.H;
    if (cin == 1 && g.p == 5) launch_k(conv0_lean_fwd_kernel<1, 5>, grid, 256, smw, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_rows);
    else if (cin == 1) launch_k(conv0_lean_fwd_kernel<1, 2>, grid, 256, smw, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_rows);
    else if (g.p == 5) launch_k(conv0_lean_fwd_kernel<2, 5>, grid, 256, smw, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_rows);
    else launch_k(conv0_lean_fwd_kernel<2, 2>, grid, 256, smw, st, x, w, bias, stat, out, out_hi, out_lo, argw, g, n_rows);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// backward of the block: d(gamma), d(beta), dW and the (zero) conv-bias gradient from dA, the winner bytes, the input
// rows and the patch moments of the forward pass; `part` needs conv0_lean_bwd_part_floats() floats
long conv0_lean_bwd_part_floats(int cin, int C, int batch, int H) {
    const long gpi = (H + kC0Rows - 1) / kC0Rows;
    return batch * gpi * (9L * cin + 1) * C + 2L * (9 * cin + 1) * C + 64;
}
int conv0_lean_backward(const float* x, int cin, int batch, const unsigned* argw, const float* dA, const PoolGeom& g,
                        const double* gram, const float* w, const float* bias, const float* gamma, const float* stat,
                        float* part, float* dw, float* db, float* dgamma, float* dbeta, cudaStream_t st) {
    const int gpi = (g.H + kC0Rows - 1) / kC0Rows, n_groups = gpi * batch, K0 = 9 * cin;
    const dim3 grid(std::min(n_groups, 2 * sm_count()), g.C / 128);
    const size_t bsm = conv0_lean_bwd_smem(cin, g.W);
    const void* kfn = cin == 1 ? (g.p == 5 ? (const void*)conv0_lean_bwd_kernel<1, 5> : (const void*)conv0_lean_bwd_kernel<1, 2>)
                               : (g.p == 5 ? (const void*)conv0_lean_bwd_kernel<2, 5> : (const void*)conv0_lean_bwd_kernel<2, 2>);
    const int rc = ensure_dyn_smem(kfn, (int)bsm);
    if (rc) return rc;
    if (cin == 1 && g.p == 5) launch_k(conv0_lean_bwd_kernel<1, 5>, grid, 256, bsm, st, x, argw, dA, g, gpi, n_groups, part);
    else if (cin == 1) launch_k(conv0_lean_bwd_kernel<1, 2>, grid, 256, bsm, st, x, argw, dA, g, gpi, n_groups, part);
    else if (g.p == 5) launch_k(conv0_lean_bwd_kernel<2, 5>, grid, 256, bsm, st, x, argw, dA, g, gpi, n_groups, part);
    else launch_k(conv0_lean_bwd_kernel<2, 2>, grid, 256, bsm, st, x, argw, dA, g, gpi, n_groups, part);
    SED_POST_LAUNCH();
    double* Ssum = reinterpret_cast<double*>(part + (size_t)grid.x * (K0 + 1) * g.C + 64);   // behind the partials
    launch_k(conv0_lean_bwd_colsum_kernel, dim3(K0 + 1, g.C / 128), 1024, 0, st, part, (int)grid.x, K0 + 1, g.C, Ssum);
    SED_POST_LAUNCH();
    launch_k(conv0_lean_bwd_finalize_kernel, (g.C * 32 + 255) / 256, 256, 0, st, Ssum, cin, g.C, gram, w, bias, gamma, stat, dw,
                                                                          db, dgamma, dbeta);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace sedb200
