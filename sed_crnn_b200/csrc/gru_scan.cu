// gru_scan.cu -- persistent recurrent scans of the bidirectional GRU.
//
// Reference: nn.GRU(bidirectional=True, batch_first=True), crnn_lightning.py:61-62,71 / sed.py:101,111
//   r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)
//   h' = (1 - z) * n + z * h ,  gh = W_hh h + b_hh ,  h_0 = 0
//
// One CTA owns (direction, tile of kBT batch rows) for all T steps: W_hh stays resident in shared
// memory (transposed so that the 3H row-dot-products read it conflict-free), the hidden state lives
// in shared memory, and gi for step t+1 is prefetched into registers while step t computes.
#include "gru_scan.cuh"

#include <cuda_bf16.h>

#include <algorithm>

namespace sedb200 {
namespace {

constexpr int kBT = 4;      // batch rows per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void gru_scan_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whh,
                                    const float* __restrict__ bhh, float* __restrict__ out,
                                    float* __restrict__ gates, int B, int T, int H) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int H3 = 3 * H;
    float* Wt = sm;                       // [H][3H]   Wt[k*3H + r] = W[r][k]
    float* h_s = Wt + (size_t)H * H3;     // [kBT][H]
    float* gh_s = h_s + kBT * H;          // [kBT][3H]
    const int dir = blockIdx.y, b0 = blockIdx.x * kBT, tid = threadIdx.x;
    const float* W = whh + (size_t)dir * H3 * H;
    for (int i = tid; i < H3 * H; i += blockDim.x) {
        const int r = i / H, k = i - r * H;
        Wt[k * H3 + r] = __ldg(W + i);
    }
    for (int i = tid; i < kBT * H; i += blockDim.x) h_s[i] = 0.0f;
    const float bias = tid < H3 ? __ldg(bhh + dir * H3 + tid) : 0.0f;
    // gate-phase role of this thread
    const int gi_item = tid / H, gj = tid - gi_item * H;
    const bool gate_thread = tid < kBT * H && (b0 + gi_item) < B;
    const long gb = b0 + gi_item;
    float nr = 0, nz = 0, nn = 0;
    auto load_gi = [&](int t) {
        const float* g = gi + ((gb * T + t) * 2 + dir) * H3;
        nr = __ldg(g + gj); nz = __ldg(g + H + gj); nn = __ldg(g + 2 * H + gj);
    };
    if (gate_thread) load_gi(dir ? T - 1 : 0);
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = dir ? T - 1 - step : step;
        const float cr = nr, cz = nz, cn = nn;
        if (gate_thread && step + 1 < T) load_gi(dir ? t - 1 : t + 1);
        // phase 1: gh[item][r] = b_hh[r] + sum_k W[r][k] h[item][k]
        if (tid < H3) {
            float acc[kBT];
#pragma unroll
            for (int i = 0; i < kBT; ++i) acc[i] = bias;
            for (int k = 0; k < H; k += 4) {
                const float w0 = Wt[(k + 0) * H3 + tid], w1 = Wt[(k + 1) * H3 + tid];
                const float w2 = Wt[(k + 2) * H3 + tid], w3 = Wt[(k + 3) * H3 + tid];
#pragma unroll
                for (int i = 0; i < kBT; ++i) {
                    const float4 hv = *reinterpret_cast<const float4*>(h_s + i * H + k);
                    acc[i] = fmaf(w0, hv.x, acc[i]);
                    acc[i] = fmaf(w1, hv.y, acc[i]);
                    acc[i] = fmaf(w2, hv.z, acc[i]);
                    acc[i] = fmaf(w3, hv.w, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < kBT; ++i) gh_s[i * H3 + tid] = acc[i];
        }
        __syncthreads();
        // phase 2: gates and the new hidden state
        if (gate_thread) {
            const float* gh = gh_s + gi_item * H3;
            const float r = sigmoidf_(cr + gh[gj]);
            const float z = sigmoidf_(cz + gh[H + gj]);
            const float q = gh[2 * H + gj];
            const float n = tanhf(fmaf(r, q, cn));
            const float hp = h_s[gi_item * H + gj];
            const float hn = fmaf(z, hp - n, n);                 // (1-z)*n + z*h
            h_s[gi_item * H + gj] = hn;
            out[(gb * T + t) * 2 * H + dir * H + gj] = hn;
            float* gs = gates + ((gb * T + t) * 2 + dir) * 4 * H;
            gs[gj] = r; gs[H + gj] = z; gs[2 * H + gj] = n; gs[3 * H + gj] = q;
        }
        __syncthreads();
    }
}

__global__ void gru_scan_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                    const float* __restrict__ gates, const float* __restrict__ whh,
                                    float* __restrict__ dgi, float* __restrict__ dgh, int B, int T, int H) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int H3 = 3 * H;
    float* W_s = sm;                      // [3H][H]
    float* dg_s = W_s + (size_t)H3 * H;   // [kBT][3H]
    const int dir = blockIdx.y, b0 = blockIdx.x * kBT, tid = threadIdx.x;
    const float* W = whh + (size_t)dir * H3 * H;
    for (int i = tid; i < H3 * H; i += blockDim.x) W_s[i] = __ldg(W + i);
    const int item = tid / H, j = tid - item * H;
    const bool active = tid < kBT * H && (b0 + item) < B;
    const long b = b0 + item;
    float dh = 0.0f;
    float n_do = 0, n_r = 0, n_z = 0, n_n = 0, n_q = 0, n_hp = 0;
    auto load_step = [&](int t) {
        n_do = __ldg(dout + (b * T + t) * 2 * H + dir * H + j);
        const float* gs = gates + ((b * T + t) * 2 + dir) * 4 * H;
        n_r = __ldg(gs + j); n_z = __ldg(gs + H + j); n_n = __ldg(gs + 2 * H + j); n_q = __ldg(gs + 3 * H + j);
        const int tp = dir ? t + 1 : t - 1;
        n_hp = (tp >= 0 && tp < T) ? __ldg(out + (b * T + tp) * 2 * H + dir * H + j) : 0.0f;
    };
    if (active) load_step(dir ? 0 : T - 1);
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = dir ? step : T - 1 - step;             // reverse of the forward order
        const float c_do = n_do, r = n_r, z = n_z, n = n_n, q = n_q, hp = n_hp;
        if (active && step + 1 < T) load_step(dir ? t + 1 : t - 1);
        float direct = 0.0f;
        if (active) {
            const float dht = c_do + dh;
            const float dn = dht * (1.0f - z);
            const float dz = dht * (hp - n);
            direct = dht * z;
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * q * r * (1.0f - r);
            const float daz = dz * z * (1.0f - z);
            const float dq = dan * r;
            const long o = ((b * T + t) * 2 + dir) * H3;
            dgi[o + j] = dar; dgi[o + H + j] = daz; dgi[o + 2 * H + j] = dan;
            dgh[o + j] = dar; dgh[o + H + j] = daz; dgh[o + 2 * H + j] = dq;
            float* ds = dg_s + item * H3;
            ds[j] = dar; ds[H + j] = daz; ds[2 * H + j] = dq;
        }
        __syncthreads();
        if (active) {
            const float* ds = dg_s + item * H3;
            float acc = direct;
            for (int rr = 0; rr < H3; rr += 4) {
                const float4 dv = *reinterpret_cast<const float4*>(ds + rr);
                acc = fmaf(W_s[(rr + 0) * H + j], dv.x, acc);
                acc = fmaf(W_s[(rr + 1) * H + j], dv.y, acc);
                acc = fmaf(W_s[(rr + 2) * H + j], dv.z, acc);
                acc = fmaf(W_s[(rr + 3) * H + j], dv.w, acc);
            }
            dh = acc;
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------ H <= 32: warp-resident scans
// One sub-warp of H lanes owns one (direction, batch row): lane j keeps row j of each W_hh gate block
// (3H registers) and its own h[j]; the matvec broadcasts h[k] with warp shuffles, so a step has no
// shared memory and no block barrier -- only the dependent FMA chains.  B*2 independent sub-warps run in
// parallel; gi for the next step is prefetched into registers.
constexpr int kCh = 8;          // time steps of operands held in registers ahead of the recurrence (forward)
constexpr int kChB = 8;         // same, backward scan
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

template <int H>
__global__ void __launch_bounds__(128)
gru_scan_fwd_warp_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                         float* __restrict__ out, float* __restrict__ gates, int B, int T) {
    pdl_wait();
    constexpr int IPW = 32 / H, H3 = 3 * H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y, j = lane % H;
    const long b = ((long)blockIdx.x * 4 + warp) * IPW + lane / H;
    const bool act = b < B;
    const float* W = whh + (size_t)dir * H3 * H;
    float w[3][H], bias[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bias[g] = __ldg(bhh + dir * H3 + g * H + j);
#pragma unroll
        for (int k = 0; k < H; ++k) w[g][k] = __ldg(W + (g * H + j) * H + k);
    }
    // gi is consumed kCh steps at a time from registers while the next kCh steps are already in flight
    // (a one-step prefetch leaves every step exposed to a full L2/HBM round trip)
    float h = 0.0f;
    float cur[kCh][3], nxt[kCh][3];
    auto load_chunk = [&](float (&dst)[kCh][3], int step0) {
#pragma unroll
        for (int s = 0; s < kCh; ++s) {
            const int step = step0 + s;
            if (act && step < T) {
                const int t = dir ? T - 1 - step : step;
                const float* g = gi + ((b * T + t) * 2 + dir) * H3;
                dst[s][0] = __ldg(g + j); dst[s][1] = __ldg(g + H + j); dst[s][2] = __ldg(g + 2 * H + j);
            } else {
                dst[s][0] = dst[s][1] = dst[s][2] = 0.0f;
            }
        }
    };
    load_chunk(cur, 0);
    for (int step0 = 0; step0 < T; step0 += kCh) {
        load_chunk(nxt, step0 + kCh);
#pragma unroll
        for (int s = 0; s < kCh; ++s) {
            const int step = step0 + s;
            if (step >= T) break;
            const int t = dir ? T - 1 - step : step;
            float a0 = bias[0], a1 = bias[1], a2 = bias[2];
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float hk = __shfl_sync(0xffffffffu, h, k, H);
                a0 = fmaf(w[0][k], hk, a0);
                a1 = fmaf(w[1][k], hk, a1);
                a2 = fmaf(w[2][k], hk, a2);
            }
            const float r = fast_sigmoid(cur[s][0] + a0);
            const float z = fast_sigmoid(cur[s][1] + a1);
            const float n = fast_tanh(fmaf(r, a2, cur[s][2]));
            h = fmaf(z, h - n, n);
            if (act) {
                out[(b * T + t) * 2 * H + dir * H + j] = h;
                float* gs = gates + ((b * T + t) * 2 + dir) * 4 * H;
                gs[j] = r; gs[H + j] = z; gs[2 * H + j] = n; gs[3 * H + j] = a2;
            }
        }
#pragma unroll
        for (int s = 0; s < kCh; ++s) { cur[s][0] = nxt[s][0]; cur[s][1] = nxt[s][1]; cur[s][2] = nxt[s][2]; }
    }
}

// Backward scan.  Besides dgi / dgh it accumulates both bias gradients per (batch row, direction) in
// registers (summed over B afterwards in a fixed order).  dW_hh = sum_(b,t) dgh (x) h_{t-1} has no
// sequential dependency and is left to a tensor-core GEMM.  W_hh columns are read from shared memory
// (conflict-free, lane j <-> column j).
template <int H>
__global__ void __launch_bounds__(128)
gru_scan_bwd_warp_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                         const float* __restrict__ gates, const float* __restrict__ whh, float* __restrict__ dgi,
                         float* __restrict__ dgh, float* __restrict__ part_w, float* __restrict__ part_b, int B,
                         int T) {
    pdl_wait();
    constexpr int IPW = 32 / H, H3 = 3 * H;
    __shared__ float Ws[H3 * H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y, j = lane % H;
    const long b = ((long)blockIdx.x * 4 + warp) * IPW + lane / H;
    const bool act = b < B;
    const float* W = whh + (size_t)dir * H3 * H;
    for (int i = threadIdx.x; i < H3 * H; i += blockDim.x) Ws[i] = __ldg(W + i);
    __syncthreads();
    float sb_r = 0, sb_z = 0, sb_n = 0, sb_q = 0;     // bias-gradient sums over t
    float dh = 0.0f;
    float cur[kChB][6], nxt[kChB][6];                   // dout, r, z, n, q, h_prev
    auto load_chunk = [&](float (&dst)[kChB][6], int step0) {
#pragma unroll
        for (int s = 0; s < kChB; ++s) {
            const int step = step0 + s;
            if (act && step < T) {
                const int t = dir ? step : T - 1 - step;
                dst[s][0] = __ldg(dout + (b * T + t) * 2 * H + dir * H + j);
                const float* gs = gates + ((b * T + t) * 2 + dir) * 4 * H;
                dst[s][1] = __ldg(gs + j); dst[s][2] = __ldg(gs + H + j);
                dst[s][3] = __ldg(gs + 2 * H + j); dst[s][4] = __ldg(gs + 3 * H + j);
                const int tp = dir ? t + 1 : t - 1;
                dst[s][5] = (tp >= 0 && tp < T) ? __ldg(out + (b * T + tp) * 2 * H + dir * H + j) : 0.0f;
            } else {
#pragma unroll
                for (int q = 0; q < 6; ++q) dst[s][q] = 0.0f;
            }
        }
    };
    load_chunk(cur, 0);
    for (int step0 = 0; step0 < T; step0 += kChB) {
        load_chunk(nxt, step0 + kChB);
#pragma unroll
        for (int s = 0; s < kChB; ++s) {
            const int step = step0 + s;
            if (step >= T) break;
            const int t = dir ? step : T - 1 - step;             // reverse of the forward order
            const float c_do = cur[s][0], r = cur[s][1], z = cur[s][2], n = cur[s][3], q = cur[s][4], hp = cur[s][5];
            const float dht = c_do + dh;
            const float dn = dht * (1.0f - z);
            const float dz = dht * (hp - n);
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * q * r * (1.0f - r);
            const float daz = dz * z * (1.0f - z);
            const float dq = dan * r;
            if (act) {
                const long o = ((b * T + t) * 2 + dir) * H3;
                dgi[o + j] = dar; dgi[o + H + j] = daz; dgi[o + 2 * H + j] = dan;
                dgh[o + j] = dar; dgh[o + H + j] = daz; dgh[o + 2 * H + j] = dq;
            }
            sb_r += dar; sb_z += daz; sb_n += dan; sb_q += dq;
            float a0 = dht * z, a1 = 0.0f, a2 = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
#pragma unroll
            for (int k = 0; k < H; k += 2) {
                a0 = fmaf(Ws[k * H + j], __shfl_sync(0xffffffffu, dar, k, H), a0);
                a1 = fmaf(Ws[(H + k) * H + j], __shfl_sync(0xffffffffu, daz, k, H), a1);
                a2 = fmaf(Ws[(2 * H + k) * H + j], __shfl_sync(0xffffffffu, dq, k, H), a2);
                b0 = fmaf(Ws[(k + 1) * H + j], __shfl_sync(0xffffffffu, dar, k + 1, H), b0);
                b1 = fmaf(Ws[(H + k + 1) * H + j], __shfl_sync(0xffffffffu, daz, k + 1, H), b1);
                b2 = fmaf(Ws[(2 * H + k + 1) * H + j], __shfl_sync(0xffffffffu, dq, k + 1, H), b2);
            }
            dh = (a0 + b0) + (a1 + b1) + (a2 + b2);
        }
#pragma unroll
        for (int s = 0; s < kChB; ++s)
#pragma unroll
            for (int q = 0; q < 6; ++q) cur[s][q] = nxt[s][q];
    }
    if (act) {
        float* pb = part_b + (b * 2) * 2 * H3;                        // [B][ih|hh][2][3H]
        pb[dir * H3 + j] = sb_r; pb[dir * H3 + H + j] = sb_z; pb[dir * H3 + 2 * H + j] = sb_n;
        pb[2 * H3 + dir * H3 + j] = sb_r; pb[2 * H3 + dir * H3 + H + j] = sb_z; pb[2 * H3 + dir * H3 + 2 * H + j] = sb_q;
    }
}


// ------------------------------------------------------------------------------ H = 32: broadcast-through-smem scans
// Same ownership as the warp-resident kernels (one warp per (direction, batch row), lane j <-> hidden unit j, its
// W_hh rows / columns in registers), but the vector every lane needs (h, or the three gate gradients) is written
// to a double-buffered shared row and read back as broadcast LDS.128 -- 8 loads instead of 32 dependent-latency
// shuffles per 32 values -- and the dot products run as packed fma.f32x2 on (k, k+1) pairs in two / four
// independent chains.  With one warp per SM sub-partition the step time IS the dependent-instruction latency,
// so this is what sets the speed of the whole scan.
constexpr int kChF32 = 8, kChB32 = 6;

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the tensor-core operand format (tc_gemm.cu)
__device__ __forceinline__ void store_plane(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long i, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// Producer / consumer hand-off between a scan warp and its helper warp through hardware named barriers (PTX
// bar.arrive + bar.sync, 64 participants): a blocked warp costs no issue slots, unlike a spin on a flag.
// None of these carries a "memory" clobber on purpose: volatile asm statements keep their order among themselves
// (barrier -> slot stores -> barrier), which is all the hand-off needs, while the compiler stays free to schedule the
// recurrence's own loads and FMAs around them -- with a clobber the hand-off would sit in the critical path.
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id)); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id)); }
__device__ __forceinline__ void ring_st(float* p, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v));
}
__device__ __forceinline__ float ring_ld(const float* p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

// REV is the scan direction as a compile-time constant: every per-step address is then `base + immediate`, and the
// integer work per step drops to two pointer bumps per chunk.
template <bool REV>
__device__ __forceinline__ void gru_fwd_bcast32_body(const float* __restrict__ gi, const float* __restrict__ whh,
                                                     const float* __restrict__ bhh, int T, long b, float (*h_s)[32],
                                                     float (*ring)[32], int ready_id, int free_id) {
    constexpr int H = 32, H3 = 96, dir = REV ? 1 : 0;
    constexpr long sGi = 2 * H3;                                           // floats per time step
    constexpr long dGi = REV ? -sGi : sGi;
    const int j = threadIdx.x & 31;
    const float* W = whh + (size_t)dir * H3 * H;
    float2 w2[3][H / 2];
    float bias[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bias[g] = __ldg(bhh + dir * H3 + g * H + j);
#pragma unroll
        for (int k = 0; k < H / 2; ++k) w2[g][k] = __ldg(reinterpret_cast<const float2*>(W + (g * H + j) * H) + k);
    }
    const long t0 = REV ? T - 1 : 0;
    const float* gp = gi + ((b * T + t0) * 2 + dir) * H3 + j;               // this lane's gi of the current chunk
    float h = 0.0f;
    float cur[kChF32][3], nxt[kChF32][3];
    auto load_chunk = [&](float (&dst)[kChF32][3], const float* g, int step0) {
#pragma unroll
        for (int s = 0; s < kChF32; ++s) {
            if (step0 + s < T) {
                dst[s][0] = __ldg(g + s * dGi); dst[s][1] = __ldg(g + s * dGi + H); dst[s][2] = __ldg(g + s * dGi + 2 * H);
            } else {
                dst[s][0] = dst[s][1] = dst[s][2] = 0.0f;
            }
        }
    };
    load_chunk(cur, gp, 0);
    int buf = 0;
    for (int step0 = 0; step0 < T; step0 += kChF32) {
        gp += kChF32 * dGi;
        load_chunk(nxt, gp, step0 + kChF32);
#pragma unroll
        for (int s = 0; s < kChF32; ++s) {
            if (step0 + s >= T) break;
            h_s[buf * 4][j] = h;
            __syncwarp();
            const float4* hv4 = reinterpret_cast<const float4*>(h_s[buf * 4]);
            buf ^= 1;
            float2 acc[3][2];
#pragma unroll
            for (int g = 0; g < 3; ++g) acc[g][0] = acc[g][1] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < H / 4; ++i) {
                const float4 hv = hv4[i];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    acc[g][0] = __ffma2_rn(w2[g][2 * i], make_float2(hv.x, hv.y), acc[g][0]);
                    acc[g][1] = __ffma2_rn(w2[g][2 * i + 1], make_float2(hv.z, hv.w), acc[g][1]);
                }
            }
            const float a0 = bias[0] + ((acc[0][0].x + acc[0][0].y) + (acc[0][1].x + acc[0][1].y));
            const float a1 = bias[1] + ((acc[1][0].x + acc[1][0].y) + (acc[1][1].x + acc[1][1].y));
            const float a2 = bias[2] + ((acc[2][0].x + acc[2][0].y) + (acc[2][1].x + acc[2][1].y));
            const float r = fast_sigmoid(cur[s][0] + a0);
            const float z = fast_sigmoid(cur[s][1] + a1);
            const float n = fast_tanh(fmaf(r, a2, cur[s][2]));
            h = fmaf(z, h - n, n);
            // hand the step's results to the helper warp, which does every global store of the scan
            if (step0 + s > 0) bar_sync64(free_id);
            ring_st(&ring[0][j], r); ring_st(&ring[1][j], z); ring_st(&ring[2][j], n); ring_st(&ring[3][j], a2);
            ring_st(&ring[4][j], h);
            bar_arrive64(ready_id);
        }
#pragma unroll
        for (int s = 0; s < kChF32; ++s) { cur[s][0] = nxt[s][0]; cur[s][1] = nxt[s][1]; cur[s][2] = nxt[s][2]; }
    }
}

// Helper warp of the forward scan: takes (r, z, n, q, h) of each step from the pair's shared-memory slot and writes
// the layer output, the saved gates and (optionally) the h_{t-1} planes -- the scan warp itself never touches HBM
// on the output side, so its dependent-instruction chain is only the recurrence.
template <bool REV>
__device__ __forceinline__ void gru_fwd_bcast32_helper(float* __restrict__ out, float* __restrict__ gates, int T, long b,
                                                       const float (*ring)[32], int ready_id, int free_id,
                                                       __nv_bfloat16* __restrict__ hp_hi, __nv_bfloat16* __restrict__ hp_lo) {
    constexpr int H = 32, dir = REV ? 1 : 0;
    constexpr long sOut = 2 * H, sGs = 2 * 4 * H;
    constexpr long dOut = REV ? -sOut : sOut, dGs = REV ? -sGs : sGs;
    const int j = threadIdx.x & 31;
    const long t0 = REV ? T - 1 : 0;
    float* op = out + (b * T + t0) * sOut + dir * H + j;
    float* sp = gates + ((b * T + t0) * 2 + dir) * 4 * H + j;
    // h_prev planes: h_t belongs to the row of the step that CONSUMES it (t+1 forward, t-1 reverse); the first
    // step's own row holds h_0 = 0
    long hrow = (b * T + t0) * 2 * H + dir * H + j;
    if (hp_hi) store_plane(hp_hi, hp_lo, hrow, 0.0f);
    for (int step = 0; step < T; ++step) {
        bar_sync64(ready_id);
        const float r = ring_ld(&ring[0][j]), z = ring_ld(&ring[1][j]), n = ring_ld(&ring[2][j]);
        const float q = ring_ld(&ring[3][j]), h = ring_ld(&ring[4][j]);
        if (step + 1 < T) bar_arrive64(free_id);
        *op = h;
        sp[0] = r; sp[H] = z; sp[2 * H] = n; sp[3 * H] = q;
        op += dOut; sp += dGs; hrow += dOut;
        if (hp_hi && step + 1 < T) store_plane(hp_hi, hp_lo, hrow, h);
    }
}

__global__ void __launch_bounds__(256)
gru_scan_fwd_bcast32_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                            float* __restrict__ out, float* __restrict__ gates, int B, int T,
                            __nv_bfloat16* __restrict__ hp_hi, __nv_bfloat16* __restrict__ hp_lo) {
    pdl_wait();
    __shared__ __align__(16) float h_s[2 * 4][32];          // [buffer][pair] rows; a pair uses rows pair and 4 + pair
    __shared__ float ring[4][5][32];                        // per pair: r, z, n, q, h of the step being handed over
    const int warp = threadIdx.x >> 5, pair = warp & 3;
    const bool helper = warp >= 4;
    const long b = (long)blockIdx.x * 4 + pair;
    if (b >= B) return;                                     // pairs are independent: no block-level barrier below
    const int ready_id = 1 + 2 * pair, free_id = 2 + 2 * pair;
    if (!helper) {
        if (blockIdx.y == 0) gru_fwd_bcast32_body<false>(gi, whh, bhh, T, b, h_s + pair, ring[pair], ready_id, free_id);
        else gru_fwd_bcast32_body<true>(gi, whh, bhh, T, b, h_s + pair, ring[pair], ready_id, free_id);
    } else {
        if (blockIdx.y == 0) gru_fwd_bcast32_helper<false>(out, gates, T, b, ring[pair], ready_id, free_id, hp_hi, hp_lo);
        else gru_fwd_bcast32_helper<true>(out, gates, T, b, ring[pair], ready_id, free_id, hp_hi, hp_lo);
    }
}

template <bool REV>
__device__ __forceinline__ void gru_bwd_bcast32_body(const float* __restrict__ dout, const float* __restrict__ out,
                                                     const float* __restrict__ gates, const float* __restrict__ whh,
                                                     float* __restrict__ part_b, int T, long b, float (*dg_s)[96],
                                                     float (*ring)[32], int ready_id, int free_id) {
    // REV = the FORWARD direction of this GRU half; the backward scan walks time the other way
    constexpr int H = 32, H3 = 96, dir = REV ? 1 : 0;
    constexpr long sOut = 2 * H, sGs = 2 * 4 * H;
    constexpr long dOut = REV ? sOut : -sOut, dGs = REV ? sGs : -sGs;
    const int j = threadIdx.x & 31;
    const float* W = whh + (size_t)dir * H3 * H;
    float2 wc[H3 / 2];                                          // (W_hh[e][j], W_hh[e+1][j]) for even e
#pragma unroll
    for (int e = 0; e < H3 / 2; ++e) wc[e] = make_float2(__ldg(W + (2 * e) * H + j), __ldg(W + (2 * e + 1) * H + j));
    const long t0 = REV ? 0 : T - 1;
    const float* dp = dout + (b * T + t0) * sOut + dir * H + j;
    const float* hp_ = out + (b * T + t0) * sOut + dir * H + j;         // h_prev of step t is out[t -/+ 1]
    const float* gsp = gates + ((b * T + t0) * 2 + dir) * 4 * H + j;
    float sb_r = 0, sb_z = 0, sb_n = 0, sb_q = 0;
    float dh = 0.0f;
    float cur[kChB32][6], nxt[kChB32][6];                       // dout, r, z, n, q, h_prev
    auto load_chunk = [&](float (&dst)[kChB32][6], const float* d, const float* hpp, const float* gs, int step0) {
#pragma unroll
        for (int s = 0; s < kChB32; ++s) {
            if (step0 + s < T) {
                dst[s][0] = __ldg(d + s * dOut);
                dst[s][1] = __ldg(gs + s * dGs); dst[s][2] = __ldg(gs + s * dGs + H);
                dst[s][3] = __ldg(gs + s * dGs + 2 * H); dst[s][4] = __ldg(gs + s * dGs + 3 * H);
                dst[s][5] = (step0 + s + 1 < T) ? __ldg(hpp + (s + 1) * dOut) : 0.0f;   // forward-previous step
            } else {
#pragma unroll
                for (int q = 0; q < 6; ++q) dst[s][q] = 0.0f;
            }
        }
    };
    load_chunk(cur, dp, hp_, gsp, 0);
    int buf = 0;
    for (int step0 = 0; step0 < T; step0 += kChB32) {
        dp += kChB32 * dOut; hp_ += kChB32 * dOut; gsp += kChB32 * dGs;
        load_chunk(nxt, dp, hp_, gsp, step0 + kChB32);
#pragma unroll
        for (int s = 0; s < kChB32; ++s) {
            if (step0 + s >= T) break;
            const float c_do = cur[s][0], r = cur[s][1], z = cur[s][2], n = cur[s][3], q = cur[s][4], hp = cur[s][5];
            const float dht = c_do + dh;
            const float dn = dht * (1.0f - z);
            const float dz = dht * (hp - n);
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * q * r * (1.0f - r);
            const float daz = dz * z * (1.0f - z);
            const float dq = dan * r;
            float* ds = dg_s[buf * 4];
            ds[j] = dar; ds[H + j] = daz; ds[2 * H + j] = dq;
            __syncwarp();
            // the helper warp turns these four values into the dgi / dgh stores (fp32 or bf16 planes)
            if (step0 + s > 0) bar_sync64(free_id);
            ring_st(&ring[0][j], dar); ring_st(&ring[1][j], daz); ring_st(&ring[2][j], dan); ring_st(&ring[3][j], dq);
            bar_arrive64(ready_id);
            sb_r += dar; sb_z += daz; sb_n += dan; sb_q += dq;
            const float4* dv4 = reinterpret_cast<const float4*>(ds);
            buf ^= 1;
            float2 acc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < H3 / 4; ++i) {
                const float4 dv = dv4[i];
                acc[(2 * i) & 3] = __ffma2_rn(wc[2 * i], make_float2(dv.x, dv.y), acc[(2 * i) & 3]);
                acc[(2 * i + 1) & 3] = __ffma2_rn(wc[2 * i + 1], make_float2(dv.z, dv.w), acc[(2 * i + 1) & 3]);
            }
            dh = fmaf(dht, z, ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + ((acc[2].x + acc[2].y) + (acc[3].x + acc[3].y)));
        }
#pragma unroll
        for (int s = 0; s < kChB32; ++s)
#pragma unroll
            for (int q = 0; q < 6; ++q) cur[s][q] = nxt[s][q];
    }
    float* pb = part_b + (b * 2) * 2 * H3;                        // [B][ih|hh][2][3H]
    pb[dir * H3 + j] = sb_r; pb[dir * H3 + H + j] = sb_z; pb[dir * H3 + 2 * H + j] = sb_n;
    pb[2 * H3 + dir * H3 + j] = sb_r; pb[2 * H3 + dir * H3 + H + j] = sb_z; pb[2 * H3 + dir * H3 + 2 * H + j] = sb_q;
}

// Helper warp of the backward scan: dgi = (dar, daz, dan), dgh = (dar, daz, dq) of each step, as fp32 or as the bf16
// hi/lo planes the tensor-core GEMMs consume.
template <bool REV>
__device__ __forceinline__ void gru_bwd_bcast32_helper(float* __restrict__ dgi, float* __restrict__ dgh, int T, long b,
                                                       const float (*ring)[32], int ready_id, int free_id,
                                                       __nv_bfloat16* __restrict__ gi_hi, __nv_bfloat16* __restrict__ gi_lo,
                                                       __nv_bfloat16* __restrict__ gh_hi, __nv_bfloat16* __restrict__ gh_lo) {
    constexpr int H = 32, H3 = 96, dir = REV ? 1 : 0;
    constexpr long sDg = 2 * H3, dDg = REV ? sDg : -sDg;
    const int j = threadIdx.x & 31;
    const long t0 = REV ? 0 : T - 1;
    long o = ((b * T + t0) * 2 + dir) * H3 + j;                 // element index in dgi / dgh and in their planes
    for (int step = 0; step < T; ++step) {
        bar_sync64(ready_id);
        const float dar = ring_ld(&ring[0][j]), daz = ring_ld(&ring[1][j]), dan = ring_ld(&ring[2][j]);
        const float dq = ring_ld(&ring[3][j]);
        if (step + 1 < T) bar_arrive64(free_id);
        if (gi_hi) {
            store_plane(gi_hi, gi_lo, o, dar); store_plane(gi_hi, gi_lo, o + H, daz); store_plane(gi_hi, gi_lo, o + 2 * H, dan);
            store_plane(gh_hi, gh_lo, o, dar); store_plane(gh_hi, gh_lo, o + H, daz); store_plane(gh_hi, gh_lo, o + 2 * H, dq);
        } else {
            dgi[o] = dar; dgi[o + H] = daz; dgi[o + 2 * H] = dan;
            dgh[o] = dar; dgh[o + H] = daz; dgh[o + 2 * H] = dq;
        }
        o += dDg;
    }
}

__global__ void __launch_bounds__(256)
gru_scan_bwd_bcast32_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                            const float* __restrict__ gates, const float* __restrict__ whh, float* __restrict__ dgi,
                            float* __restrict__ dgh, float* __restrict__ part_b, int B, int T,
                            __nv_bfloat16* __restrict__ gi_hi, __nv_bfloat16* __restrict__ gi_lo,
                            __nv_bfloat16* __restrict__ gh_hi, __nv_bfloat16* __restrict__ gh_lo) {
    pdl_wait();
    __shared__ __align__(16) float dg_s[2 * 4][96];
    __shared__ float ring[4][4][32];                        // per pair: dar, daz, dan, dq of the step being handed over
    const int warp = threadIdx.x >> 5, pair = warp & 3;
    const bool helper = warp >= 4;
    const long b = (long)blockIdx.x * 4 + pair;
    if (b >= B) return;
    const int ready_id = 1 + 2 * pair, free_id = 2 + 2 * pair;
    if (!helper) {
        if (blockIdx.y == 0) gru_bwd_bcast32_body<false>(dout, out, gates, whh, part_b, T, b, dg_s + pair, ring[pair], ready_id, free_id);
        else gru_bwd_bcast32_body<true>(dout, out, gates, whh, part_b, T, b, dg_s + pair, ring[pair], ready_id, free_id);
    } else {
        if (blockIdx.y == 0) gru_bwd_bcast32_helper<false>(dgi, dgh, T, b, ring[pair], ready_id, free_id, gi_hi, gi_lo, gh_hi, gh_lo);
        else gru_bwd_bcast32_helper<true>(dgi, dgh, T, b, ring[pair], ready_id, free_id, gi_hi, gi_lo, gh_hi, gh_lo);
    }
}

// ------------------------------------------------------------------------------ H = 64 / 128: K-split scans
// W_hh no longer fits one lane per unit, so each hidden unit is served by FOUR adjacent lanes that each keep a
// quarter of the unit's three W_hh rows in registers (3*H/4 values) and reduce with two xor-shuffles; the CTA
// (4*H threads) owns kSplitBT batch rows of one direction, the hidden state is double-buffered in shared memory
// (one block barrier per step).  Lane `ks` of a unit's quad also does the gate math of batch row `ks`.
constexpr int kSplitBT = 2;

template <int H>
__global__ void __launch_bounds__(4 * H, 1)
gru_scan_fwd_split_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                          float* __restrict__ out, float* __restrict__ gates, int B, int T) {
    pdl_wait();
    constexpr int KS = H / 4, H3 = 3 * H, HP = H + 16;          // HP: row pitch with 4 floats of padding per K-slice
    __shared__ __align__(16) float h_s[2][kSplitBT][HP];
    const int tid = threadIdx.x, ks = tid & 3, j = tid >> 2;
    const int dir = blockIdx.y, b0 = blockIdx.x * kSplitBT;
    const float* W = whh + (size_t)dir * H3 * H;
    float w[3][KS];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) w[g][kk] = __ldg(W + (size_t)(g * H + j) * H + ks * KS + kk);
    float bias[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) bias[g] = __ldg(bhh + dir * H3 + g * H + j);
    for (int i = tid; i < 2 * kSplitBT * HP; i += blockDim.x) (&h_s[0][0][0])[i] = 0.0f;
    const bool gate = ks < kSplitBT && (b0 + ks) < B;            // this lane finishes unit j of batch row b0 + ks
    const long b = b0 + ks;
    float hprev = 0.0f, nr = 0, nz = 0, nn = 0;
    auto load_gi = [&](int t) {
        const float* g = gi + ((b * T + t) * 2 + dir) * H3;
        nr = __ldg(g + j); nz = __ldg(g + H + j); nn = __ldg(g + 2 * H + j);
    };
    if (gate) load_gi(dir ? T - 1 : 0);
    __syncthreads();
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir ? T - 1 - step : step;
        const float cr = nr, cz = nz, cn = nn;
        if (gate && step + 1 < T) load_gi(dir ? t - 1 : t + 1);
        float acc[kSplitBT][3];
#pragma unroll
        for (int r = 0; r < kSplitBT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < KS; kk += 4) {
#pragma unroll
            for (int r = 0; r < kSplitBT; ++r) {
                const float4 hv = *reinterpret_cast<const float4*>(&h_s[cur][r][ks * (KS + 4) + kk]);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    acc[r][g] = fmaf(w[g][kk + 0], hv.x, acc[r][g]);
                    acc[r][g] = fmaf(w[g][kk + 1], hv.y, acc[r][g]);
                    acc[r][g] = fmaf(w[g][kk + 2], hv.z, acc[r][g]);
                    acc[r][g] = fmaf(w[g][kk + 3], hv.w, acc[r][g]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kSplitBT; ++r)
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                acc[r][g] += __shfl_xor_sync(0xffffffffu, acc[r][g], 1);
                acc[r][g] += __shfl_xor_sync(0xffffffffu, acc[r][g], 2);
            }
        if (gate) {
            const float g0 = (ks == 0 ? acc[0][0] : acc[1][0]) + bias[0];
            const float g1 = (ks == 0 ? acc[0][1] : acc[1][1]) + bias[1];
            const float g2 = (ks == 0 ? acc[0][2] : acc[1][2]) + bias[2];
            const float r = fast_sigmoid(cr + g0);
            const float z = fast_sigmoid(cz + g1);
            const float n = fast_tanh(fmaf(r, g2, cn));
            hprev = fmaf(z, hprev - n, n);
            h_s[cur ^ 1][ks][j + 4 * (j / KS)] = hprev;
            out[(b * T + t) * 2 * H + dir * H + j] = hprev;
            float* gs = gates + ((b * T + t) * 2 + dir) * 4 * H;
            gs[j] = r; gs[H + j] = z; gs[2 * H + j] = n; gs[3 * H + j] = g2;
        }
        __syncthreads();
        cur ^= 1;
    }
}

template <int H>
__global__ void __launch_bounds__(4 * H, 1)
gru_scan_bwd_split_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                          const float* __restrict__ gates, const float* __restrict__ whh, float* __restrict__ dgi,
                          float* __restrict__ dgh, float* __restrict__ part_b, int B, int T) {
    pdl_wait();
    constexpr int H3 = 3 * H, RS = H3 / 4, DP = H3 + 16;        // RS rows of W_hh per K-slice; padded pitch
    __shared__ __align__(16) float dg_s[2][kSplitBT][DP];
    const int tid = threadIdx.x, ks = tid & 3, j = tid >> 2;
    const int dir = blockIdx.y, b0 = blockIdx.x * kSplitBT;
    const float* W = whh + (size_t)dir * H3 * H;
    float wc[RS];                                               // W_hh[ks*RS + i][j]
#pragma unroll
    for (int i = 0; i < RS; ++i) wc[i] = __ldg(W + (size_t)(ks * RS + i) * H + j);
    const bool gate = ks < kSplitBT && (b0 + ks) < B;
    const long b = b0 + ks;
    float dh = 0.0f, sb_r = 0, sb_z = 0, sb_n = 0, sb_q = 0;
    float n_do = 0, n_r = 0, n_z = 0, n_n = 0, n_q = 0, n_hp = 0;
    auto load_step = [&](int t) {
        n_do = __ldg(dout + (b * T + t) * 2 * H + dir * H + j);
        const float* gs = gates + ((b * T + t) * 2 + dir) * 4 * H;
        n_r = __ldg(gs + j); n_z = __ldg(gs + H + j); n_n = __ldg(gs + 2 * H + j); n_q = __ldg(gs + 3 * H + j);
        const int tp = dir ? t + 1 : t - 1;
        n_hp = (tp >= 0 && tp < T) ? __ldg(out + (b * T + tp) * 2 * H + dir * H + j) : 0.0f;
    };
    if (gate) load_step(dir ? 0 : T - 1);
    for (int i = tid; i < 2 * kSplitBT * DP; i += blockDim.x) (&dg_s[0][0][0])[i] = 0.0f;
    __syncthreads();
    // slot of gradient element e (0..3H-1) inside the padded row: 4 floats of padding per K-slice
    auto pad = [&](int e) { return e + 4 * (e / RS); };
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir ? step : T - 1 - step;
        float direct = 0.0f;
        if (gate) {
            const float c_do = n_do, r = n_r, z = n_z, n = n_n, q = n_q, hp = n_hp;
            if (step + 1 < T) load_step(dir ? t + 1 : t - 1);
            const float dht = c_do + dh;
            const float dn = dht * (1.0f - z);
            const float dz = dht * (hp - n);
            direct = dht * z;
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * q * r * (1.0f - r);
            const float daz = dz * z * (1.0f - z);
            const float dq = dan * r;
            const long o = ((b * T + t) * 2 + dir) * H3;
            dgi[o + j] = dar; dgi[o + H + j] = daz; dgi[o + 2 * H + j] = dan;
            dgh[o + j] = dar; dgh[o + H + j] = daz; dgh[o + 2 * H + j] = dq;
            sb_r += dar; sb_z += daz; sb_n += dan; sb_q += dq;
            dg_s[cur][ks][pad(j)] = dar;
            dg_s[cur][ks][pad(H + j)] = daz;
            dg_s[cur][ks][pad(2 * H + j)] = dq;
        }
        __syncthreads();
        float acc[kSplitBT];
#pragma unroll
        for (int r = 0; r < kSplitBT; ++r) acc[r] = 0.0f;
#pragma unroll
        for (int i = 0; i < RS; i += 4) {
#pragma unroll
            for (int r = 0; r < kSplitBT; ++r) {
                const float4 dv = *reinterpret_cast<const float4*>(&dg_s[cur][r][ks * (RS + 4) + i]);
                acc[r] = fmaf(wc[i + 0], dv.x, acc[r]);
                acc[r] = fmaf(wc[i + 1], dv.y, acc[r]);
                acc[r] = fmaf(wc[i + 2], dv.z, acc[r]);
                acc[r] = fmaf(wc[i + 3], dv.w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kSplitBT; ++r) {
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 2);
        }
        if (gate) dh = direct + (ks == 0 ? acc[0] : acc[1]);
        cur ^= 1;
    }
    if (gate) {
        float* pb = part_b + (b * 2) * 2 * H3;                  // [B][ih|hh][2][3H]
        pb[dir * H3 + j] = sb_r; pb[dir * H3 + H + j] = sb_z; pb[dir * H3 + 2 * H + j] = sb_n;
        pb[2 * H3 + dir * H3 + j] = sb_r; pb[2 * H3 + dir * H3 + H + j] = sb_z; pb[2 * H3 + dir * H3 + 2 * H + j] = sb_q;
    }
}

// ------------------------------------------------------------------------------ H = 128: register-tiled backward scan
constexpr int kOpDepth = 6;             // cp.async ring slots (operands are requested kOpDepth - 1 steps ahead)
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const float* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The K-split kernel above reads 96 gradient words per thread per row and step: 393 KB of shared-memory traffic per
// step and SM, which is what bounds it (3,072 LSU cycles per step against 768 FMA-pipe cycles).  Here a thread owns an
// 8 (hidden units) x 12 (gate-gradient slice) tile of W_hh -- still 96 registers -- so each loaded word feeds 8
// FMAs: 6 LDS.128 per step instead of 48.  The two batch rows of a block are interleaved in shared memory and travel
// together through packed fma.f32x2 (weight as the scalar operand), and the 32 slices of a hidden unit are folded
// with a reduce-scatter over the warp (18 shuffles per step).  Direction is a template constant.
template <bool REV>
__device__ __forceinline__ void gru_bwd_tile128_body(const float* __restrict__ dout, const float* __restrict__ out,
                                                     const float* __restrict__ gates, const float* __restrict__ whh,
                                                     float* __restrict__ dgi, float* __restrict__ dgh,
                                                     float* __restrict__ part_b, int B, int T, float (*dg_s)[384][2],
                                                     float (*ops)[6][128][2]) {
    constexpr int H = 128, H3 = 384, dir = REV ? 1 : 0;
    constexpr long sOut = 2 * H, sGs = 2 * 4 * H, sDg = 2 * H3;           // floats per time step
    constexpr long dOut = REV ? sOut : -sOut, dGs = REV ? sGs : -sGs, dDg = REV ? sDg : -sDg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* W = whh + (size_t)dir * H3 * H;
    float w[8][12];                                           // W_hh[12*lane + ee][8*warp + jj]
#pragma unroll
    for (int ee = 0; ee < 12; ++ee)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) w[jj][ee] = __ldg(W + (size_t)(12 * lane + ee) * H + 8 * warp + jj);
    // after the reduce-scatter lane L holds hidden unit 8*warp + ((L >> 2) & 7) of both rows; lanes with (L & 3) < 2
    // finish (row L & 3, that unit): gate derivatives, outputs, next step's shared vector
    const int j = 8 * warp + ((lane >> 2) & 7), r = lane & 3;
    const long b = (long)blockIdx.x * 2 + r;
    const bool gate = r < 2 && b < B;
    const long t0 = REV ? 0 : T - 1;
    // element indices as 32-bit integers (the launcher checks the tensors have < 2^31 elements): one register each
    // instead of five 64-bit pointers -- the kernel sits exactly at the 128-register limit of a 512-thread block
    int io = (int)((b * T + t0) * sOut + dir * H + j);                     // into dout / out
    int ig = (int)(((b * T + t0) * 2 + dir) * 4 * H + j);                  // into gates
    int id = (int)(((b * T + t0) * 2 + dir) * H3 + j);                     // into dgi / dgh
    float dh = 0.0f, sb_r = 0, sb_z = 0, sb_n = 0, sb_q = 0;
    // The six per-step operands of a gate thread travel global -> shared with cp.async, kOpDepth - 1 steps ahead of
    // their use: no registers are held for data in flight (there are none to spare), and a step no longer waits a
    // DRAM round trip for operands requested only one step earlier.
    float* my_ops = &ops[0][0][(lane >> 2) + 8 * warp][r & 1];          // [slot][operand][unit][row]
    const unsigned ops_addr = (unsigned)__cvta_generic_to_shared(my_ops);
    constexpr unsigned kOpStride = 128 * 2 * 4, kSlotStride = 6 * kOpStride;     // bytes
    auto issue = [&](int step) {                              // operands of `step`; io / ig point at that step
        if (gate && step < T) {
            const unsigned d = ops_addr + (unsigned)(step % kOpDepth) * kSlotStride;
            cp_async4(d, dout + io);
            cp_async4(d + kOpStride, gates + ig);
            cp_async4(d + 2 * kOpStride, gates + ig + H);
            cp_async4(d + 3 * kOpStride, gates + ig + 2 * H);
            cp_async4(d + 4 * kOpStride, gates + ig + 3 * H);
            if (step + 1 < T) cp_async4(d + 5 * kOpStride, out + io + dOut);     // forward-previous hidden state
            io += (int)dOut; ig += (int)dGs;
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int p = 0; p < kOpDepth - 1; ++p) issue(p);
    for (int i = threadIdx.x; i < 2 * H3 * 2; i += blockDim.x) (&dg_s[0][0][0])[i] = 0.0f;
    __syncthreads();
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        issue(step + kOpDepth - 1);
        cp_async_wait<kOpDepth - 1>();                        // this step's operands have landed
        float direct = 0.0f;
        if (gate) {
            const float* o = my_ops + (size_t)(step % kOpDepth) * (kSlotStride / 4);
            const float c_do = o[0], rr = o[kOpStride / 4], z = o[2 * (kOpStride / 4)], n = o[3 * (kOpStride / 4)];
            const float q = o[4 * (kOpStride / 4)], hp = step + 1 < T ? o[5 * (kOpStride / 4)] : 0.0f;
            const float dht = c_do + dh;
            const float dn = dht * (1.0f - z);
            const float dz = dht * (hp - n);
            direct = dht * z;
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * q * rr * (1.0f - rr);
            const float daz = dz * z * (1.0f - z);
            const float dq = dan * rr;
            dgi[id] = dar; dgi[id + H] = daz; dgi[id + 2 * H] = dan;
            dgh[id] = dar; dgh[id + H] = daz; dgh[id + 2 * H] = dq;
            id += (int)dDg;
            sb_r += dar; sb_z += daz; sb_n += dan; sb_q += dq;
            dg_s[cur][j][r] = dar;
            dg_s[cur][H + j][r] = daz;
            dg_s[cur][2 * H + j][r] = dq;
        }
        __syncthreads();
        // dh[j] (both rows) += sum over this lane's 12 gradient words
        float2 acc[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[jj] = make_float2(0.0f, 0.0f);
        const float4* dv = reinterpret_cast<const float4*>(&dg_s[cur][12 * lane][0]);
#pragma unroll
        for (int q4 = 0; q4 < 6; ++q4) {
            const float4 d4 = dv[q4];                          // (e, row0), (e, row1), (e+1, row0), (e+1, row1)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                acc[jj] = __ffma2_rn(make_float2(w[jj][2 * q4], w[jj][2 * q4]), make_float2(d4.x, d4.y), acc[jj]);
                acc[jj] = __ffma2_rn(make_float2(w[jj][2 * q4 + 1], w[jj][2 * q4 + 1]), make_float2(d4.z, d4.w), acc[jj]);
            }
        }
        // reduce-scatter over the 32 slices: 8 -> 4 -> 2 -> 1 values per lane, then two butterflies
#pragma unroll
        for (int half = 4, off = 16; half >= 1; half >>= 1, off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float2 send = up ? acc[i] : acc[i + half], keep = up ? acc[i + half] : acc[i];
                acc[i].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, off);
                acc[i].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, off);
            }
        }
        acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, 2); acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, 2);
        acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, 1); acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, 1);
        if (gate) dh = direct + (r == 0 ? acc[0].x : acc[0].y);
        cur ^= 1;
    }
    if (gate) {
        float* pb = part_b + (b * 2) * 2 * H3;                  // [B][ih|hh][2][3H]
        pb[dir * H3 + j] = sb_r; pb[dir * H3 + H + j] = sb_z; pb[dir * H3 + 2 * H + j] = sb_n;
        pb[2 * H3 + dir * H3 + j] = sb_r; pb[2 * H3 + dir * H3 + H + j] = sb_z; pb[2 * H3 + dir * H3 + 2 * H + j] = sb_q;
    }
}

__global__ void __launch_bounds__(512, 1)
gru_scan_bwd_tile128_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                            const float* __restrict__ gates, const float* __restrict__ whh, float* __restrict__ dgi,
                            float* __restrict__ dgh, float* __restrict__ part_b, int B, int T) {
    pdl_wait();
    __shared__ __align__(16) float dg_s[2][384][2];           // [buffer][gradient word][row]
    __shared__ float ops[kOpDepth][6][128][2];                // cp.async ring of the gate threads' operands
    if (blockIdx.y == 0) gru_bwd_tile128_body<false>(dout, out, gates, whh, dgi, dgh, part_b, B, T, dg_s, ops);
    else gru_bwd_tile128_body<true>(dout, out, gates, whh, dgi, dgh, part_b, B, T, dg_s, ops);
}

// ------------------------------------------------------------------------------ H = 128: register-tiled forward scan
// Thread tile: 3 gates x 2 hidden units x 16 k (96 weights): 8 LDS.128 per step instead of 16, rows paired in
// fma.f32x2, the 8 k-slices of a unit pair folded with a reduce-scatter (18 shuffles), gi through a cp.async ring.
template <bool REV>
__device__ __forceinline__ void gru_fwd_tile128_body(const float* __restrict__ gi, const float* __restrict__ whh,
                                                     const float* __restrict__ bhh, float* __restrict__ out,
                                                     float* __restrict__ gates, int B, int T, float (*h_s)[128][2],
                                                     float (*ops)[3][128][2]) {
    constexpr int H = 128, H3 = 384, dir = REV ? 1 : 0;
    constexpr long sGi = 2 * H3, sOut = 2 * H, sGs = 2 * 4 * H;
    constexpr long dGi = REV ? -sGi : sGi, dOut = REV ? -sOut : sOut, dGs = REV ? -sGs : sGs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int up = lane >> 3, ks = lane & 7;                  // unit pair of the warp, k slice
    const int j0 = 8 * warp + 2 * up;
    const float* W = whh + (size_t)dir * H3 * H;
    float w[3][2][16];                                        // W_hh[g*H + j0 + u][k], k = 2*(ks + 8*i) + {0, 1}
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(W + (size_t)(g * H + j0 + u) * H + 2 * (ks + 8 * i)));
                w[g][u][2 * i] = t.x;
                w[g][u][2 * i + 1] = t.y;
            }
    // after the reduce-scatter, lanes with bit 2 == u hold the three gate sums of unit j0 + u (both rows); of those,
    // lanes with (lane & 3) < 2 finish (row lane & 3, that unit)
    const int j = j0 + ((lane >> 2) & 1), r = lane & 3;
    const long b = (long)blockIdx.x * 2 + r;
    const bool gate = r < 2 && b < B;
    const long t0 = REV ? T - 1 : 0;
    int ig = (int)(((b * T + t0) * 2 + dir) * H3 + j);        // into gi
    int io = (int)((b * T + t0) * sOut + dir * H + j);        // into out
    int is = (int)(((b * T + t0) * 2 + dir) * 4 * H + j);     // into gates
    float bias[3] = {0.f, 0.f, 0.f};
    if (gate) {
#pragma unroll
        for (int g = 0; g < 3; ++g) bias[g] = __ldg(bhh + dir * H3 + g * H + j);
    }
    float* my_ops = &ops[0][0][j][r & 1];
    const unsigned ops_addr = (unsigned)__cvta_generic_to_shared(my_ops);
    constexpr unsigned kOpStride = 128 * 2 * 4, kSlotStride = 3 * kOpStride;
    auto issue = [&](int step) {
        if (gate && step < T) {
            const unsigned d = ops_addr + (unsigned)(step % kOpDepth) * kSlotStride;
            cp_async4(d, gi + ig);
            cp_async4(d + kOpStride, gi + ig + H);
            cp_async4(d + 2 * kOpStride, gi + ig + 2 * H);
            ig += (int)dGi;
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int p = 0; p < kOpDepth - 1; ++p) issue(p);
    for (int i = threadIdx.x; i < 2 * H * 2; i += blockDim.x) (&h_s[0][0][0])[i] = 0.0f;
    float hprev = 0.0f;
    __syncthreads();
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        issue(step + kOpDepth - 1);
        float2 acc[3][2];
#pragma unroll
        for (int g = 0; g < 3; ++g) acc[g][0] = acc[g][1] = make_float2(0.0f, 0.0f);
        const float4* hv = reinterpret_cast<const float4*>(&h_s[cur][0][0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 h4 = hv[ks + 8 * i];                 // (k, row0), (k, row1), (k+1, row0), (k+1, row1)
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    acc[g][u] = __ffma2_rn(make_float2(w[g][u][2 * i], w[g][u][2 * i]), make_float2(h4.x, h4.y), acc[g][u]);
                    acc[g][u] = __ffma2_rn(make_float2(w[g][u][2 * i + 1], w[g][u][2 * i + 1]), make_float2(h4.z, h4.w), acc[g][u]);
                }
        }
        // fold the 8 k slices: lanes with bit 2 clear keep unit 0, the others unit 1; then two butterflies
        float2 v[3];
        {
            const bool upper = (lane & 4) != 0;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float2 send = upper ? acc[g][0] : acc[g][1], keep = upper ? acc[g][1] : acc[g][0];
                v[g].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 4);
                v[g].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 4);
            }
#pragma unroll
            for (int off = 2; off >= 1; off >>= 1)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    v[g].x += __shfl_xor_sync(0xffffffffu, v[g].x, off);
                    v[g].y += __shfl_xor_sync(0xffffffffu, v[g].y, off);
                }
        }
        cp_async_wait<kOpDepth - 1>();                        // this step's gi has landed
        if (gate) {
            const float* o = my_ops + (size_t)(step % kOpDepth) * (kSlotStride / 4);
            const float g0 = (r == 0 ? v[0].x : v[0].y) + bias[0];
            const float g1 = (r == 0 ? v[1].x : v[1].y) + bias[1];
            const float g2 = (r == 0 ? v[2].x : v[2].y) + bias[2];
            const float rr = fast_sigmoid(o[0] + g0);
            const float z = fast_sigmoid(o[kOpStride / 4] + g1);
            const float n = fast_tanh(fmaf(rr, g2, o[2 * (kOpStride / 4)]));
            hprev = fmaf(z, hprev - n, n);
            h_s[cur ^ 1][j][r] = hprev;
            out[io] = hprev;
            gates[is] = rr; gates[is + H] = z; gates[is + 2 * H] = n; gates[is + 3 * H] = g2;
            io += (int)dOut; is += (int)dGs;
        }
        __syncthreads();
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(512, 1)
gru_scan_fwd_tile128_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                            float* __restrict__ out, float* __restrict__ gates, int B, int T) {
    pdl_wait();
    __shared__ __align__(16) float h_s[2][128][2];            // [buffer][hidden unit][row]
    __shared__ float ops[kOpDepth][3][128][2];                // cp.async ring of gi
    if (blockIdx.y == 0) gru_fwd_tile128_body<false>(gi, whh, bhh, out, gates, B, T, h_s, ops);
    else gru_fwd_tile128_body<true>(gi, whh, bhh, out, gates, B, T, h_s, ops);
}

template <int H>
int launch_warp_fwd(const float* gi, const float* whh, const float* bhh, float* out, float* gates, int B, int T,
                    cudaStream_t st) {
    constexpr int per_block = 4 * (32 / H);
    dim3 grid((B + per_block - 1) / per_block, 2);
    launch_k(gru_scan_fwd_warp_kernel<H>, grid, 128, 0, st, gi, whh, bhh, out, gates, B, T);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}
template <int H>
int launch_warp_bwd(const float* dout, const float* out, const float* gates, const float* whh, float* dgi,
                    float* dgh, float* part_w, float* part_b, int B, int T, cudaStream_t st) {
    constexpr int per_block = 4 * (32 / H);
    dim3 grid((B + per_block - 1) / per_block, 2);
    launch_k(gru_scan_bwd_warp_kernel<H>, grid, 128, 0, st, dout, out, gates, whh, dgi, dgh, part_w, part_b, B, T);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

inline int round32(int v) { return (v + 31) / 32 * 32; }

}  // namespace

bool gru_scan_emits_planes(int H) { return H == 32; }

int gru_scan_forward(const float* gi, const float* whh, const float* bhh, float* out, float* gates, int B, int T,
                     int H, cudaStream_t st, void* hprev_hi, void* hprev_lo) {
    SED_REQUIRE(!hprev_hi || gru_scan_emits_planes(H), SEDB200_ESHAPE, "gru_scan: h_prev planes need H = 32 (H = %d)", H);
    if (H == 32) {
        launch_k(gru_scan_fwd_bcast32_kernel, dim3((B + 3) / 4, 2), 256, 0, st,
            gi, whh, bhh, out, gates, B, T, reinterpret_cast<__nv_bfloat16*>(hprev_hi), reinterpret_cast<__nv_bfloat16*>(hprev_lo));
        SED_POST_LAUNCH();
        return SEDB200_OK;
    }
    if (H == 16) return launch_warp_fwd<16>(gi, whh, bhh, out, gates, B, T, st);
    if (H == 8) return launch_warp_fwd<8>(gi, whh, bhh, out, gates, B, T, st);
    if (H == 128 || H == 64) {
        SED_REQUIRE((long)B * T * 2 * 4 * H < (1L << 31), SEDB200_ESHAPE, "gru_scan: B*T too large for H=%d (B=%d T=%d)", H, B, T);
        dim3 grid((B + kSplitBT - 1) / kSplitBT, 2);
        if (H == 128) launch_k(gru_scan_fwd_tile128_kernel, grid, 512, 0, st, gi, whh, bhh, out, gates, B, T);
        else launch_k(gru_scan_fwd_split_kernel<64>, grid, 256, 0, st, gi, whh, bhh, out, gates, B, T);
        SED_POST_LAUNCH();
        return SEDB200_OK;
    }
    const int threads = round32(std::max(3 * H, kBT * H));
    SED_REQUIRE(threads <= 1024 && H % 4 == 0, SEDB200_ESHAPE, "gru_scan: H=%d unsupported", H);
    const size_t smem = ((size_t)3 * H * H + kBT * H + kBT * 3 * H) * 4;
    { const int rc = ensure_dyn_smem((const void*)gru_scan_fwd_kernel, 227 * 1024); if (rc) return rc; }
    SED_REQUIRE(smem <= 227 * 1024, SEDB200_ESHAPE, "gru_scan: H=%d needs %zu B shared memory", H, smem);
    dim3 grid((B + kBT - 1) / kBT, 2);
    launch_k(gru_scan_fwd_kernel, grid, threads, smem, st, gi, whh, bhh, out, gates, B, T, H);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

bool gru_scan_fused_param_grads(int H) { return H == 128 || H == 64 || H == 32 || H == 16 || H == 8; }

int gru_scan_backward(const float* dout, const float* out, const float* gates, const float* whh, float* dgi,
                      float* dgh, float* part_w, float* part_b, int B, int T, int H, cudaStream_t st,
                      void* const* planes) {
    SED_REQUIRE(!planes || gru_scan_emits_planes(H), SEDB200_ESHAPE, "gru_scan: gradient planes need H = 32 (H = %d)", H);
    if (H == 32) {
        __nv_bfloat16* pl[4] = {nullptr, nullptr, nullptr, nullptr};
        if (planes) for (int i = 0; i < 4; ++i) pl[i] = reinterpret_cast<__nv_bfloat16*>(planes[i]);
        launch_k(gru_scan_bwd_bcast32_kernel, dim3((B + 3) / 4, 2), 256, 0, st, dout, out, gates, whh, dgi, dgh, part_b, B, T,
                                                                           pl[0], pl[1], pl[2], pl[3]);
        SED_POST_LAUNCH();
        return SEDB200_OK;
    }
    if (H == 16) return launch_warp_bwd<16>(dout, out, gates, whh, dgi, dgh, part_w, part_b, B, T, st);
    if (H == 8) return launch_warp_bwd<8>(dout, out, gates, whh, dgi, dgh, part_w, part_b, B, T, st);
    if (H == 128 || H == 64) {
        SED_REQUIRE((long)B * T * 2 * 4 * H < (1L << 31), SEDB200_ESHAPE, "gru_scan: B*T too large for H=%d (B=%d T=%d)", H, B, T);
        dim3 grid((B + kSplitBT - 1) / kSplitBT, 2);
        if (H == 128) launch_k(gru_scan_bwd_tile128_kernel, grid, 512, 0, st, dout, out, gates, whh, dgi, dgh, part_b, B, T);
        else launch_k(gru_scan_bwd_split_kernel<64>, grid, 256, 0, st, dout, out, gates, whh, dgi, dgh, part_b, B, T);
        SED_POST_LAUNCH();
        return SEDB200_OK;
    }
    const int threads = round32(kBT * H);
    SED_REQUIRE(threads <= 1024 && H % 4 == 0, SEDB200_ESHAPE, "gru_scan: H=%d unsupported", H);
    const size_t smem = ((size_t)3 * H * H + kBT * 3 * H) * 4;
    { const int rc = ensure_dyn_smem((const void*)gru_scan_bwd_kernel, 227 * 1024); if (rc) return rc; }
    SED_REQUIRE(smem <= 227 * 1024, SEDB200_ESHAPE, "gru_scan: H=%d needs %zu B shared memory", H, smem);
    dim3 grid((B + kBT - 1) / kBT, 2);
    launch_k(gru_scan_bwd_kernel, grid, threads, smem, st, dout, out, gates, whh, dgi, dgh, B, T, H);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace sedb200

using namespace sedb200;

extern "C" {

int sedb200_gru_scan_fused_bias_grads(int H) { return gru_scan_fused_param_grads(H) ? 1 : 0; }

int sedb200_gru_scan_fwd(const float* gi_dev, const float* whh_dev, const float* bhh_dev, float* out_dev,
                         float* gates_dev, int B, int T, int H, void* stream) {
    SED_REQUIRE(gi_dev && whh_dev && bhh_dev && out_dev && gates_dev, SEDB200_EINVAL, "gru_scan_fwd: null buffer");
    SED_REQUIRE(B >= 1 && T >= 1 && H >= 1, SEDB200_EINVAL, "gru_scan_fwd: B=%d T=%d H=%d", B, T, H);
    int rc = require_sm100();
    if (rc) return rc;
    return gru_scan_forward(gi_dev, whh_dev, bhh_dev, out_dev, gates_dev, B, T, H, as_stream(stream));
}

int sedb200_gru_scan_bwd(const float* dout_dev, const float* out_dev, const float* gates_dev, const float* whh_dev,
                         float* dgi_dev, float* dgh_dev, float* part_b_dev, int B, int T, int H, void* stream) {
    SED_REQUIRE(dout_dev && out_dev && gates_dev && whh_dev && dgi_dev && dgh_dev, SEDB200_EINVAL, "gru_scan_bwd: null buffer");
    SED_REQUIRE(B >= 1 && T >= 1 && H >= 1, SEDB200_EINVAL, "gru_scan_bwd: B=%d T=%d H=%d", B, T, H);
    SED_REQUIRE(part_b_dev || !gru_scan_fused_param_grads(H), SEDB200_EINVAL,
                "gru_scan_bwd: H=%d writes bias-gradient partials, part_b_dev must be given", H);
    int rc = require_sm100();
    if (rc) return rc;
    return gru_scan_backward(dout_dev, out_dev, gates_dev, whh_dev, dgi_dev, dgh_dev, nullptr, part_b_dev, B, T, H,
                             as_stream(stream));
}

}  // extern "C"
