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
// One CTA per SM: blockIdx.y = 64 channels = two groups of 32 (B stays resident: 80 KB), blockIdx.x walks window tiles.
// Warp roles: warps 0-3 build A rows (thread <-> window) into a double-buffered, 128-byte-swizzled K-major tile;
// warp 4 issues the MMAs (3-term bf16 split, one TMEM accumulator per channel group: the epilogue frees it as soon as
// the tile sits in its registers); warps 8-23 are the epilogue (group, TMEM sub-partition = warp % 4, channel half).
// Measured (ncu, profiles/): the kernel is bound by instruction issue -- ~40 instructions per (window, channel) for
// max / argmax / ReLU / dropout / fp16 hi-lo packing / winner bytes -- not by the tensor pipe (15 %) or HBM.
#include <cuda_bf16.h>
#include "conv0_lean.cuh"
#include "tc_umma.cuh"

#include <algorithm>

namespace sedb200 {
namespace {
using namespace umma;

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
// 32 lanes x 8 consecutive fp32 columns, no wait
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int CIN, int P>
struct WinDims {
    static constexpr int PW = P + 2;                              // patch width
    static constexpr int KP = CIN * 3 * PW;                       // patch entries
    static constexpr int KV = KP + 1;                             // + the constant-1 column (bias / shift)
    static constexpr int KSTEPS = (KV + 15) / 16;
    static constexpr int NB = 32 * P;                             // B rows of one channel group = MMA N
    static constexpr int BPlane = NB * 128;                       // bytes
    static constexpr int kGroups = 2;                             // channel groups (of 32) per CTA
    static constexpr int kSmem = 1024 + kGroups * 2 * BPlane + 4 * kWinAPlane + 128;
};

// warps 0-3: A producers; warp 4: MMA issuer (5-7 idle); warps 8-23: epilogue (group, TMEM sub-partition, channel half)
constexpr int kWinThreads = 768;

template <int CIN, int P>
__global__ void __launch_bounds__(kWinThreads, 1)
conv0_win_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ stat, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                     __nv_bfloat16* __restrict__ out_lo, unsigned* __restrict__ argw, PoolGeom g, unsigned n_windows,
                     int n_tiles) {
    pdl_wait();
    using D = WinDims<CIN, P>;
    extern __shared__ unsigned char c0w_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c0w_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* b_base = smem;                                 // [group][hi | lo]
    unsigned char* a_base = smem + D::kGroups * 2 * D::BPlane;    // [buf][hi | lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + 4 * kWinAPlane);
    uint64_t *a_full = bars, *a_empty = bars + 2, *t_full = bars + 4, *t_empty = bars + 6;     // t_*: one per channel group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp-uniform by construction
    const int cg0 = blockIdx.y * (32 * D::kGroups);               // first channel of this CTA

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1);
            mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    // ---- B operand (resident): group gq, row n = j*32 + cl  <->  (window position j, channel cg0 + gq*32 + cl)
    for (int idx = tid; idx < D::kGroups * 2 * D::NB; idx += kWinThreads) {
        const int gq = idx / (2 * D::NB), rem = idx - gq * 2 * D::NB;
        const bool lo = rem >= D::NB;
        const int n = lo ? rem - D::NB : rem;
        const int j = n >> 5, c = cg0 + gq * 32 + (n & 31);
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
        put_row<D::KV>(b_base + (gq * 2 + (lo ? 1 : 0)) * D::BPlane, n, v, lo);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const unsigned uWo = (unsigned)g.Wo, uH = (unsigned)g.H;

    if (warp < 4) {
        // ================= A producers: thread <-> window row =================
        const int r = tid;
        int buf = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const unsigned q = (unsigned)tile * 128u + (unsigned)r;
            const bool valid = q < n_windows;
            const unsigned bh = q / uWo, wo = q - bh * uWo, b = bh / uH, h = bh - b * uH;
            const int w0 = (int)wo * P - 1;                       // first patch column; only the two edge columns can be
            const bool lok = w0 >= 0, rok = w0 + D::PW - 1 < g.W; // outside the image (W >= Wo * P)
            float v[D::KV];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int hh = (int)h + rr - 1;
                    const bool rowok = valid && hh >= 0 && hh < g.H;
                    const float* src = x + (((long)b * CIN + ci) * g.H + (rowok ? hh : 0)) * g.W + w0;
                    float* vr = v + (ci * 3 + rr) * D::PW;
                    vr[0] = (rowok && lok) ? __ldg(src) : 0.0f;
#pragma unroll
                    for (int u = 1; u < D::PW - 1; ++u) vr[u] = rowok ? __ldg(src + u) : 0.0f;
                    vr[D::PW - 1] = (rowok && rok) ? __ldg(src + D::PW - 1) : 0.0f;
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
        // ================= MMA issuer: one accumulator per channel group, released by the epilogue as soon as it
        //                   sits in registers (a few hundred cycles), so a single buffer per group is enough =========
        // (warp-wide control flow, one elected lane issues: tc_umma.cuh elect_one)
        {
            constexpr uint32_t idesc = idesc_bf16(128, D::NB, 0, 0);
            const uint64_t dbase = smem_desc_sw128(smem_u32(b_base), 16, 1024);          // b_base = start of the buffer
            constexpr uint32_t apl_u = kWinAPlane >> 4, bpl_u = D::BPlane >> 4, a_off_u = (D::kGroups * 2 * D::BPlane) >> 4;
            int buf = 0; uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(a_full + buf, phase);
                const uint64_t dah0 = dbase + (uint64_t)(a_off_u + buf * 2 * apl_u), dal0 = dah0 + apl_u;
#pragma unroll
                for (int gq = 0; gq < D::kGroups; ++gq) {
                    mbar_wait(t_empty + gq, tphase ^ 1);         // the epilogue has drained this group's accumulator
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t dbh0 = dbase + (uint64_t)(gq * 2 * bpl_u), dbl0 = dbh0 + bpl_u;
                        const uint32_t d = tmem + gq * 256;
#pragma unroll
                        for (int k = 0; k < D::KSTEPS; ++k) {
                            mma_bf16(d, dah0 + 2 * k, dbh0 + 2 * k, idesc, k != 0);
                            mma_bf16(d, dah0 + 2 * k, dbl0 + 2 * k, idesc, 1);
                            mma_bf16(d, dal0 + 2 * k, dbh0 + 2 * k, idesc, 1);
                        }
                        if (gq == D::kGroups - 1) mma_commit(a_empty + buf);
                        mma_commit(t_full + gq);
                    }
                    __syncwarp();
                }
                tphase ^= 1;
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================= epilogue: lane <-> window; warp <-> (group, sub-partition, 16 channels in two passes of 8)
        const int e = warp - 8, sp = e & 3, chh = (e >> 2) & 1, gq = e >> 3;
        const int C4 = g.C >> 2;
        const int c_first = cg0 + gq * 32 + chh * 16;             // first of this warp's 16 channels
        const float keep_scale = g.drop_p > 0.0f ? 1.0f / (1.0f - g.drop_p) : 1.0f;
        uint32_t tphase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(t_full + gq, tphase);
            tc_fence_after();
            const uint32_t ta = tmem + ((uint32_t)(sp * 32) << 16) + gq * 256 + chh * 16;
            const unsigned q = (unsigned)tile * 128u + (unsigned)(sp * 32 + lane);
            const bool live = q < n_windows;
            unsigned words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                      // two passes of 8 channels: 40 accumulator registers
                float z[P][8];
#pragma unroll
                for (int j = 0; j < P; ++j) tmem_ld8_nowait(ta + j * 32 + hf * 8, z[j]);
                tmem_ld_wait();
                if (hf == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + gq);     // registers hold the tile: the accumulator is free
                }
                if (live) {
                    float4 pend = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int mm = 0; mm < 2; ++mm) {
                        const int m = 2 * hf + mm;
                        const long i = (long)q * C4 + ((c_first >> 2) + m);   // element numbering of the pool kernels (dropout)
                        Keep4 kp;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) kp.k[qq] = true;
                        if (g.drop_p > 0.0f) kp = dropout_keep4(pool_seed(g), (unsigned long long)i, g.drop_p);
                        float mv[4];
                        unsigned word = 0;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const int k = 4 * mm + qq;
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
                        if (mm == 0) pend = m4;
                        else if (out_hi) store_planes8(out_hi, out_lo, (long)q, (c_first >> 2) + m - 1, C4, pend, m4);
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
                }
            }
            tphase ^= 1;
            if (live) {
                if (argw)
                    *reinterpret_cast<uint4*>(argw + (long)q * C4 + (c_first >> 2)) = make_uint4(words[0], words[1], words[2], words[3]);
            }
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
    const int ny = g.C / (32 * D::kGroups);
    const dim3 grid(std::max(1, std::min(n_tiles, sm_count() / ny)), ny);
    const void* kfn = (const void*)conv0_win_fwd_kernel<CIN, P>;
    const int rc = ensure_dyn_smem(kfn, D::kSmem);
    if (rc) return rc;
    launch_k(conv0_win_fwd_kernel<CIN, P>, grid, kWinThreads, D::kSmem, st, x, w, bias, stat, out, out_hi, out_lo, argw, g,
                                                                       n_windows, n_tiles);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace

bool conv0_win_ok(int cin, int C, int pool, long n_windows) {
    return (cin == 1 || cin == 2) && (pool == 5 || pool == 2) && C % 64 == 0 && C / 64 <= 64 && n_windows > 0 &&
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
