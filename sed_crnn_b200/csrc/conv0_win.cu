// conv0_win.cu -- lean block 0 forward, "window rows" formulation on tcgen05.
//
// Reference ops: Conv2d(3x3, pad 1) -> BatchNorm2d(train) -> ReLU -> MaxPool2d((1,p)) -> Dropout of the FIRST block
// (crnn_lightning.py:46-52 / sed.py:87-92,106-107).  Same contract as conv0_tc_fwd_kernel (conv0_lean.cu): pooled output
// (fp16 hi / lo planes and / or fp32) + one winner byte per (window, channel), statistics known up front.
//
// What was wrong with pixel rows: the accumulator hands every thread one PIXEL, so the max over a pooling window is a
// cross-thread operation -- the whole tile went TMEM -> registers -> shared memory -> registers again, and that tail was
// two thirds of the kernel's 100 M warp instructions (profiles/README.md, "lean block 0").  Here a row of the MMA is a
// pooling WINDOW and the window position lives in the N dimension:
//
//   A row   = the (3 x (p+2) x Cin) input patch under a window, plus a constant 1          K' = 3(p+2)Cin + 1 <= 43
//   B row n = (position j, channel c): the 3x3 kernel of channel c placed at column offset j of that patch, already
//             multiplied by the BatchNorm scale; the constant column carries scale*bias + shift
//   D[window][j*32 + c] = BatchNorm(conv)(pixel j of the window, channel c)                 N = 32 p per CTA
//
// so the thread that owns a TMEM lane holds all p positions of its window in registers: max / argmax / ReLU / dropout
// are plain register code and the stores go straight out.  The structural zeros in B and the padded K cost MMA work
// (23.6 MFLOP per 128 windows at p = 5, ~35 us per step in total) which runs beside the epilogue on the tensor pipe.
//
// One CTA per SM: blockIdx.y = a group of 32 channels (B stays resident: 40 KB), blockIdx.x walks window tiles.
// Warp roles: warps 0-3 build A rows (thread <-> window) into a double-buffered, 128-byte-swizzled K-major tile;
// warp 4 issues the MMAs (3-term bf16 split, two TMEM accumulators); warps 8-15 are the epilogue (TMEM sub-partition =
// warp % 4, channel half = (warp - 8) / 4).
#include <cuda_bf16.h>
#include "conv0_lean.cuh"
#include "tc_umma.cuh"

#include <algorithm>

namespace sedb200 {
namespace {
using namespace umma;

constexpr int kWinThreads = 512;
constexpr int kWinAPlane = 128 * 128;                             // bytes: 128 rows x 128 B (64 bf16 K slots, <= 48 used)

__device__ __forceinline__ uint32_t sw128(int r, int c) {         // byte offset of 16-byte chunk c of row r
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// one K-major row (KV valid entries, zero-padded to whole 16-wide k-steps) as a bf16 hi or lo plane
template <int KV>
__device__ __forceinline__ void put_row(unsigned char* plane, int r, const float (&v)[KV], bool lo) {
    constexpr int CH = ((KV + 15) / 16) * 2;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = c * 8 + 2 * e;
            const float a = k0 < KV ? v[k0 < KV ? k0 : 0] : 0.0f, b = k0 + 1 < KV ? v[k0 + 1 < KV ? k0 + 1 : 0] : 0.0f;
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

// 32 lanes x 16 consecutive fp32 columns, no wait (the caller waits once for all its loads)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int CIN, int P>
struct WinDims {
    static constexpr int PW = P + 2;                              // patch width
    static constexpr int KP = CIN * 3 * PW;                       // patch entries
    static constexpr int KV = KP + 1;                             // + the constant-1 column (bias / shift)
    static constexpr int KSTEPS = (KV + 15) / 16;
    static constexpr int NB = 32 * P;                             // B rows = MMA N
    static constexpr int BPlane = NB * 128;                       // bytes
    static constexpr int kSmem = 1024 + 2 * BPlane + 4 * kWinAPlane + 128;
};

template <int CIN, int P>
__global__ void __launch_bounds__(kWinThreads, 1)
conv0_win_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ stat, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                     __nv_bfloat16* __restrict__ out_lo, unsigned* __restrict__ argw, PoolGeom g, unsigned n_windows,
                     int n_tiles) {
    using D = WinDims<CIN, P>;
    extern __shared__ unsigned char c0w_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c0w_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *b_hi = smem, *b_lo = smem + D::BPlane;
    unsigned char* a_base = smem + 2 * D::BPlane;                 // [buf][hi | lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + 4 * kWinAPlane);
    uint64_t *a_full = bars, *a_empty = bars + 2, *t_full = bars + 4, *t_empty = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg0 = blockIdx.y * 32;                              // first channel of this CTA's group

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1);
            mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    // ---- B operand (resident): row n = j*32 + cl  <->  (window position j, channel cg0 + cl)
    for (int idx = tid; idx < 2 * D::NB; idx += kWinThreads) {
        const bool lo = idx >= D::NB;
        const int n = lo ? idx - D::NB : idx;
        const int j = n >> 5, c = cg0 + (n & 31);
        const float sc = __ldg(stat + 2 * g.C + c), sh = __ldg(stat + 3 * g.C + c);
        float v[D::KV];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int u = 0; u < D::PW; ++u) {
                    const int t = u - j;
                    v[(ci * 3 + rr) * D::PW + u] = (t >= 0 && t <= 2) ? sc * __ldg(w + ((long)c * CIN + ci) * 9 + rr * 3 + t) : 0.0f;
                }
        v[D::KP] = fmaf(__ldg(bias + c), sc, sh);
        put_row<D::KV>(lo ? b_lo : b_hi, n, v, lo);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const unsigned uWo = (unsigned)g.Wo, uH = (unsigned)g.H;

    if (warp < 4) {
        // ================= A producers: thread <-> window row =================
        const int r = tid;
        int buf = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const unsigned q = (unsigned)tile * 128u + (unsigned)r;
            const bool valid = q < n_windows;
            const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
            float v[D::KV];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int hh = (int)h + rr - 1;
                    const bool rowok = valid && hh >= 0 && hh < g.H;
                    const float* src = x + (((long)b * CIN + ci) * g.H + (rowok ? hh : 0)) * g.W;
#pragma unroll
                    for (int u = 0; u < D::PW; ++u) {
                        const int ww = (int)wo * P + u - 1;
                        const bool ok = rowok && ww >= 0 && ww < g.W;
                        v[(ci * 3 + rr) * D::PW + u] = ok ? __ldg(src + ww) : 0.0f;
                    }
                }
            v[D::KP] = valid ? 1.0f : 0.0f;
            mbar_wait(a_empty + buf, phase ^ 1);                 // the MMAs that read this buffer have retired
            unsigned char* ah = a_base + buf * 2 * kWinAPlane;
            put_row<D::KV>(ah, r, v, false);
            put_row<D::KV>(ah + kWinAPlane, r, v, true);
            fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core
            mbar_arrive(a_full + buf);
            if (++buf == 2) { buf = 0; phase ^= 1; }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(128, D::NB, 0, 0);
            const uint32_t bh2 = smem_u32(b_hi), bl = smem_u32(b_lo);
            int buf = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(t_empty + buf, phase ^ 1);             // the epilogue has drained this accumulator
                mbar_wait(a_full + buf, phase);
                tc_fence_after();
                const uint32_t ah = smem_u32(a_base + buf * 2 * kWinAPlane), al = ah + kWinAPlane;
                const uint32_t d = tmem + buf * 256;
#pragma unroll
                for (int k = 0; k < D::KSTEPS; ++k) {
                    const uint64_t dah = smem_desc_sw128(ah + k * 32, 16, 1024), dal = smem_desc_sw128(al + k * 32, 16, 1024);
                    const uint64_t dbh = smem_desc_sw128(bh2 + k * 32, 16, 1024), dbl = smem_desc_sw128(bl + k * 32, 16, 1024);
                    mma_bf16(d, dah, dbh, idesc, k != 0);
                    mma_bf16(d, dah, dbl, idesc, 1);
                    mma_bf16(d, dal, dbh, idesc, 1);
                }
                mma_commit(a_empty + buf);
                mma_commit(t_full + buf);
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: lane <-> window, 16 channels per warp =================
        const int e = warp - 8, sp = e & 3, chh = e >> 2;
        const int C4 = g.C >> 2;
        const int c_first = cg0 + chh * 16;                       // first of this warp's 16 channels
        const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
        int buf = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(t_full + buf, phase);
            tc_fence_after();
            float z[P][16];
            const uint32_t ta = tmem + ((uint32_t)(sp * 32) << 16) + buf * 256 + chh * 16;
#pragma unroll
            for (int j = 0; j < P; ++j) tmem_ld16_nowait(ta + j * 32, z[j]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + buf);            // registers hold the tile: the accumulator is free
            const unsigned q = (unsigned)tile * 128u + (unsigned)(sp * 32 + lane);
            if (q < n_windows) {
                unsigned words[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const long i = (long)q * C4 + ((c_first >> 2) + m);   // element numbering of the pool kernels (dropout)
                    Keep4 kp;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) kp.k[qq] = true;
                    if (g.drop_p > 0.0f) kp = dropout_keep4(g.seed, (unsigned long long)i, g.drop_p);
                    float mv[4];
                    unsigned word = 0;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int k = 4 * m + qq;
                        float best = z[0][k];
                        unsigned arg = 0u;
#pragma unroll
                        for (int j = 1; j < P; ++j)
                            if (z[j][k] > best) { best = z[j][k]; arg = (unsigned)j; }      // first maximum wins
                        const bool alive = best > 0.0f && kp.k[qq];
                        mv[qq] = alive ? best * keep_scale : 0.0f;
                        word |= (arg | (alive ? 0u : 0x80u)) << (8 * qq);
                    }
                    words[m] = word;
                    const float4 m4 = make_float4(mv[0], mv[1], mv[2], mv[3]);
                    if (out_hi) store_planes4(out_hi, out_lo, i, m4);
                    if (out) {
                        const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
                        float* dst = out + (long)b * g.oB + (long)h * g.oH + (long)wo * g.oW + (long)(c_first + 4 * m) * g.oC;
                        if (g.oC == 1) {
                            *reinterpret_cast<float4*>(dst) = m4;
                        } else {
                            dst[0] = mv[0]; dst[g.oC] = mv[1]; dst[2 * g.oC] = mv[2]; dst[3 * g.oC] = mv[3];
                        }
                    }
                }
                if (argw)
                    *reinterpret_cast<uint4*>(argw + (long)q * C4 + (c_first >> 2)) = make_uint4(words[0], words[1], words[2], words[3]);
            }
            if (++buf == 2) { buf = 0; phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int CIN, int P>
int launch_win(const float* x, int batch, const float* w, const float* bias, const float* stat, const PoolGeom& g,
               float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, unsigned* argw, cudaStream_t st) {
    using D = WinDims<CIN, P>;
    const unsigned n_windows = (unsigned)((long)batch * g.H * g.Wo);
    const int n_tiles = (int)((n_windows + 127u) / 128u);
    const int ny = g.C / 32;
    const dim3 grid(std::max(1, std::min(n_tiles, sm_count() / ny)), ny);
    const void* kfn = (const void*)conv0_win_fwd_kernel<CIN, P>;
    const int rc = ensure_dyn_smem(kfn, D::kSmem);
    if (rc) return rc;
    conv0_win_fwd_kernel<CIN, P><<<grid, kWinThreads, D::kSmem, st>>>(x, w, bias, stat, out, out_hi, out_lo, argw, g,
                                                                       n_windows, n_tiles);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}


// ====================================================================================== backward
// S[k][c] = sum over windows of dz * patch_k(winner), S[K][c] = sum dz  (what conv0_lean_bwd_kernel produces: the
// only data-dependent part of block 0's backward, conv0_lean.cu header) as a tensor-core contraction in the same
// window coordinates as the forward:
//
//   D_j[c][k'] = sum over windows q of dZ_j[q][c] * patch[q][k'],   dZ_j[q][c] = dz[q][c] if winner(q, c) == j else 0
//
// i.e. p accumulators of 128 channels x K' patch entries that stay in TMEM for the whole kernel; the fold back to the
// 3x3 taps -- S[(ci, r, t)][c] = sum_j D_j[c][(ci, r, t + j)] -- happens once per CTA at the end.  Both operands are
// "MN-major" straight from row-per-window shared tiles (the layout wgrad_tc_kernel uses): a K-block is 32 windows.
// dz and the patches are carried as bf16 hi / lo planes, three MMAs per k-step (fp32-grade, like every other
// contraction whose result feeds Adam through a BatchNorm-cancelled sum).
// Warps 0-7 build the tiles (thread <-> window, 16 channels), warp 8 issues, warps 0-3 drain TMEM at the end.
constexpr int kBwThreads = 288;
constexpr int kBwWin = 32;                                        // windows per K-block
constexpr int kBwBox = kBwWin * 128;                              // 4 KB: 32 rows x 128 B

template <int CIN, int P>
struct WinBwdDims {
    using F = WinDims<CIN, P>;
    static constexpr int NPAD = F::KSTEPS * 16;                   // MMA N: patch entries + constant, padded
    static constexpr int kDzPlane = 2 * kBwBox;                   // one position, one of hi / lo: 2 channel halves
    static constexpr int kStage = P * 2 * kDzPlane + 2 * kBwBox;  // dZ (p positions x hi, lo) + patches (hi, lo)
    static constexpr int kSmem = 1024 + 2 * kStage + 128;
};

template <int CIN, int P>
__global__ void __launch_bounds__(kBwThreads, 1)
conv0_win_bwd_kernel(const float* __restrict__ x, const unsigned* __restrict__ argw, const float* __restrict__ dA,
                     PoolGeom g, unsigned n_windows, int n_tiles, float* __restrict__ part) {
    using D = WinBwdDims<CIN, P>;
    using F = WinDims<CIN, P>;
    constexpr int K = CIN * 9;
    extern __shared__ unsigned char c0b_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c0b_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * D::kStage);
    uint64_t *full = bars, *empty = bars + 2, *done = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.y * 128;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(full + i, 256); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const unsigned uWo = (unsigned)g.Wo, uH = (unsigned)g.H;
    const int C4 = g.C >> 2;
    const bool any_tile = (int)blockIdx.x < n_tiles;

    if (warp < 8) {
        // ================= producers: thread <-> (window wl of the K-block, 16-channel slice o) =================
        const int wl = tid >> 3, o = tid & 7;
        const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
        int buf = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const unsigned q = (unsigned)tile * kBwWin + (unsigned)wl;
            const bool valid = q < n_windows;
            // ---- loads first (they overlap the wait for the buffer)
            float gq[16];
            unsigned wb[4];
            {
                const float4* src = reinterpret_cast<const float4*>(dA + (long)(valid ? q : 0) * g.C + c0 + o * 16);
                const uint4 ww = __ldg(reinterpret_cast<const uint4*>(argw + (long)(valid ? q : 0) * C4 + (c0 >> 2) + o * 4));
                wb[0] = ww.x; wb[1] = ww.y; wb[2] = ww.z; wb[3] = ww.w;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const float4 t4 = __ldg(src + v4);
                    gq[4 * v4] = t4.x; gq[4 * v4 + 1] = t4.y; gq[4 * v4 + 2] = t4.z; gq[4 * v4 + 3] = t4.w;
                }
            }
            float pv[8];                                          // patch chunk o (8 entries) of window wl
            constexpr int PCH = F::KSTEPS * 2;                    // 16-byte chunks of a patch row in use
            if (o < PCH) {
                const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = o * 8 + e;
                    float val = 0.0f;
                    if (k < F::KP) {
                        const int cr = k / F::PW, u = k - cr * F::PW, ci = cr / 3, rr = cr - ci * 3;
                        const int hh = (int)h + rr - 1, ww2 = (int)wo * P + u - 1;
                        if (valid && hh >= 0 && hh < g.H && ww2 >= 0 && ww2 < g.W)
                            val = __ldg(x + (((long)b * CIN + ci) * g.H + hh) * g.W + ww2);
                    } else if (k == F::KP) {
                        val = valid ? 1.0f : 0.0f;
                    }
                    pv[e] = val;
                }
            }
            // ---- dz as bf16 hi / lo halves, and the winner of every channel
            unsigned short hi16[16], lo16[16];
            unsigned win[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned byte = (wb[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                const float dz = (valid && !(byte & 0x80u)) ? gq[k] * keep_scale : 0.0f;
                const __nv_bfloat16 h = __float2bfloat16_rn(dz);
                hi16[k] = __bfloat16_as_ushort(h);
                lo16[k] = __bfloat16_as_ushort(__float2bfloat16_rn(dz - __bfloat162float(h)));
                win[k] = (byte & 0x80u) ? 0xFFu : (byte & 0x7Fu);
            }
            mbar_wait(empty + buf, phase ^ 1);
            unsigned char* st = smem + buf * D::kStage;
            // dZ planes: [j][hi|lo][half][32 rows x 128 B]; this thread's 16 channels = chunks (o & 3) * 2, +1 of half o >> 2
#pragma unroll
            for (int j = 0; j < P; ++j) {
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    unsigned char* base = st + (j * 2 + pl) * D::kDzPlane + (o >> 2) * kBwBox;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k0 = cc * 8 + 2 * e;
                            const unsigned a = pl ? lo16[k0] : hi16[k0], bb = pl ? lo16[k0 + 1] : hi16[k0 + 1];
                            wv[e] = (win[k0] == (unsigned)j ? a : 0u) | ((win[k0 + 1] == (unsigned)j ? bb : 0u) << 16);
                        }
                        *reinterpret_cast<uint4*>(base + sw128(wl, (o & 3) * 2 + cc)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
            if (o < PCH) {
                unsigned char* pb = st + P * 2 * D::kDzPlane;
                uint32_t wh[4], wlw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(pv[2 * e]), h1 = __float2bfloat16_rn(pv[2 * e + 1]);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(pv[2 * e] - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(pv[2 * e + 1] - __bfloat162float(h1));
                    wh[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    wlw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                *reinterpret_cast<uint4*>(pb + sw128(wl, o)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(pb + kBwBox + sw128(wl, o)) = make_uint4(wlw[0], wlw[1], wlw[2], wlw[3]);
            }
            fence_proxy_async();
            mbar_arrive(full + buf);
            if (++buf == 2) { buf = 0; phase ^= 1; }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(128, D::NPAD, 1, 1);      // both operands MN-major
            int buf = 0; uint32_t phase = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(full + buf, phase);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + buf * D::kStage);
                const uint32_t p_hi = st + P * 2 * D::kDzPlane, p_lo = p_hi + kBwBox;
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const uint32_t z_hi = st + (j * 2) * D::kDzPlane, z_lo = z_hi + D::kDzPlane;
                    const uint32_t d = tmem + j * 64;
#pragma unroll
                    for (int k = 0; k < kBwWin / 16; ++k) {
                        const uint64_t dah = smem_desc_sw128(z_hi + k * 2048, kBwBox, 1024), dal = smem_desc_sw128(z_lo + k * 2048, kBwBox, 1024);
                        const uint64_t dbh = smem_desc_sw128(p_hi + k * 2048, kBwBox, 1024), dbl = smem_desc_sw128(p_lo + k * 2048, kBwBox, 1024);
                        mma_bf16(d, dah, dbh, idesc, (first && k == 0) ? 0u : 1u);
                        mma_bf16(d, dah, dbl, idesc, 1);
                        mma_bf16(d, dal, dbh, idesc, 1);
                    }
                }
                first = false;
                mma_commit(empty + buf);
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
            mma_commit(done);
        }
    }
    // ================= drain: thread <-> channel, fold the window coordinates back to the 3x3 taps =================
    if (warp < 4) {
        float S[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) S[k] = 0.0f;
        if (any_tile) {
            mbar_wait(done, 0);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < P; ++j) {
                float dd[D::NPAD];
#pragma unroll
                for (int cc = 0; cc < D::NPAD / 16; ++cc) {
                    float t16[16];
                    tmem_ld16_nowait(tmem + ((uint32_t)(warp * 32) << 16) + j * 64 + cc * 16, t16);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) dd[cc * 16 + e] = t16[e];
                }
#pragma unroll
                for (int cr = 0; cr < CIN * 3; ++cr)
#pragma unroll
                    for (int t = 0; t < 3; ++t) S[cr * 3 + t] += dd[cr * F::PW + t + j];
                S[K] += dd[F::KP];
            }
        }
        const int c = c0 + warp * 32 + lane;
#pragma unroll
        for (int k = 0; k <= K; ++k) part[((long)blockIdx.x * (K + 1) + k) * g.C + c] = S[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int CIN, int P>
int launch_win_bwd(const float* x, int batch, const unsigned* argw, const float* dA, const PoolGeom& g, float* part,
                   int* nblk, cudaStream_t st) {
    using D = WinBwdDims<CIN, P>;
    const unsigned n_windows = (unsigned)((long)batch * g.H * g.Wo);
    const int n_tiles = (int)((n_windows + kBwWin - 1) / kBwWin);
    const int ny = g.C / 128;
    const dim3 grid(std::max(1, std::min(n_tiles, sm_count() / ny)), ny);
    const void* kfn = (const void*)conv0_win_bwd_kernel<CIN, P>;
    const int rc = ensure_dyn_smem(kfn, D::kSmem);
    if (rc) return rc;
    conv0_win_bwd_kernel<CIN, P><<<grid, kBwThreads, D::kSmem, st>>>(x, argw, dA, g, n_windows, n_tiles, part);
    SED_POST_LAUNCH();
    *nblk = (int)grid.x;
    return SEDB200_OK;
}

}  // namespace

bool conv0_win_ok(int cin, int C, int pool, long n_windows) {
    return (cin == 1 || cin == 2) && (pool == 5 || pool == 2) && C % 32 == 0 && C / 32 <= 64 && n_windows > 0 &&
           n_windows < (1L << 31);
}

int conv0_win_forward(const float* x, int cin, int batch, const float* w, const float* bias, const float* stat,
                      const PoolGeom& g, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, unsigned* argw,
                      cudaStream_t st) {
    if (cin == 1 && g.p == 5) return launch_win<1, 5>(x, batch, w, bias, stat, g, out, out_hi, out_lo, argw, st);
    if (cin == 1) return launch_win<1, 2>(x, batch, w, bias, stat, g, out, out_hi, out_lo, argw, st);
    if (g.p == 5) return launch_win<2, 5>(x, batch, w, bias, stat, g, out, out_hi, out_lo, argw, st);
    return launch_win<2, 2>(x, batch, w, bias, stat, g, out, out_hi, out_lo, argw, st);
}

}  // namespace sedb200

namespace sedb200 {

// S partials of block 0's backward on tcgen05: part[blk][9*cin + 1][C] (the layout conv0_lean_bwd_colsum_kernel reduces);
// *nblk = number of partial blocks written.  Needs the channels-last block output (g.oC == 1) and C % 128 == 0.
bool conv0_win_bwd_ok(int cin, const PoolGeom& g, long n_windows) {
    return conv0_win_ok(cin, g.C, g.p, n_windows) && g.C % 128 == 0 && g.oC == 1;
}
long conv0_win_bwd_max_blocks() { return sm_count(); }
int conv0_win_backward_partials(const float* x, int cin, int batch, const unsigned* argw, const float* dA,
                                const PoolGeom& g, float* part, int* nblk, cudaStream_t st) {
    if (cin == 1 && g.p == 5) return launch_win_bwd<1, 5>(x, batch, argw, dA, g, part, nblk, st);
    if (cin == 1) return launch_win_bwd<1, 2>(x, batch, argw, dA, g, part, nblk, st);
    if (g.p == 5) return launch_win_bwd<2, 5>(x, batch, argw, dA, g, part, nblk, st);
    return launch_win_bwd<2, 2>(x, batch, argw, dA, g, part, nblk, st);
}

}  // namespace sedb200
