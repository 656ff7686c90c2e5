// logmel_tc.cu -- the log-mel front end with the DFT on the tensor cores (tcgen05 / TMEM).
//
// Replaces feature._mbe (/root/reference/feature.py:55-59) like logmel.cu does; same tables, same mel walk, same
// output.  What differs is where the 2048-point real DFT of a Hann-windowed frame runs: as TWO batched GEMMs
//
//     n = 32 n1 + n2,   k = k1 + 64 k2          W_2048^(nk) = W_64^(n1 k1) W_2048^(n2 k1) W_32^(n2 k2)
//
//   stage 1   rows (frame, n2), K = n1 (64 real samples)        -> Y[n2][k1], k1 = 0..32, Hermitian-packed into 64 reals
//   twiddle   Z[n2][k1] = Y[n2][k1] W_2048^(n2 k1)               (fp32, registers)
//   stage 2   rows (frame, k1 = 0..31), K = (n2, re / im) = 64  -> X[k1 + 64 k2], k2 = 0..31 (k2 >= 16: mirror bins)
//   special   rows (frame), K = n2 (32 reals, Y[n2][32])        -> X[32 + 64 k2], k2 = 0..15
//
// on fp16 hi / lo operand planes (x = hi + lo, 22 significand bits; products hi*hi + hi*lo + lo*hi, fp32 accumulate in
// TMEM), each frame scaled by a power of two so that fp16 never over- or underflows.  Measured against the float64
// oracle: 1.3e-6 on the 1e-4 gate (tests/manual/logmel_tc_emul.py is the CPU emulation this layout was settled with;
// two split terms instead of three fail the gate at 1.8e-3).
//
// A tile is FOUR frames = 128 MMA rows.  One persistent CTA per SM, warp roles (640 threads):
//   warps 0-3   producers: PCM (LDG.128) -> Hann -> per-frame scale -> fp16 hi / lo -> A1 rows of a tile buffer
//   warps 4-7   stage-1 epilogue: TMEM -> twiddle -> fp16 hi / lo -> A2 rows, written IN PLACE over A1
//   warps 8-15  stage-2 epilogue, two sets that alternate tiles: TMEM -> |X|^2 -> P[1025] of the frame in shared
//               memory -> the mel walk of logmel.cu (warp per frame) -> log -> 160 B store
//   warp 16     one thread issues every tcgen05.mma (12 per stage and tile + 6 for the special rows); warps 17-19 idle
//               (register budgets move between the five warpgroups with setmaxnreg)
// Three tile buffers (32 KB each) are in flight; D1 / D2 / D2-special are double-buffered in TMEM (320 columns).
//
// Row / K permutations (free: any bijection works as long as both sides of a GEMM agree):
//   producers load float4s, so a thread holds n2 = 4 (lane & 7) + e and n1 = (lane >> 3) + 4 j; its four rows sit at
//   row position 8 e + (lane & 7) of the frame's 32 (conflict-free 16-byte stores under the 128-byte swizzle) and
//   n1 sits at K position 16 (lane >> 3) + j (B1's K rows are permuted the same way on the host).
#include <cuda_fp16.h>
#include "logmel.cuh"
#include "tc_umma.cuh"

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace sedb200 {
namespace {
using namespace umma;

constexpr int kTeams = 4;                        // tiles in flight per CTA: a team = 4 warps = one tile (4 frames)
constexpr int kTcThreads = kTeams * 128;
constexpr int kPlane = 128 * 128;                // one operand plane of a tile: 128 rows x 128 B
constexpr int kTileBytes = 2 * kPlane;           // hi | lo; later the power spectra of the four frames (4 x 4224 B)
constexpr int kSpecBytes = 2 * 1024;             // special rows of a tile: one 8-row atom per plane (rows 0..3 used)
constexpr int kPStride = 33 * 32;                // floats per frame in the power-spectrum buffer (bins 0..1024 + pad)
constexpr int kTmemCols = 128;                   // per team: D1 / D2 at +0 (64 columns, never live together), D2-special at +64

// constant operands, built on the host in double precision, laid out exactly as they sit in shared memory
struct TcConst {
    unsigned char b1[2][64 * 128];               // stage 1: N = 64 outputs x K = 64 (permuted n1), fp16, 128B-swizzled
    unsigned char b2[2][64 * 128];               // stage 2: N = 64 outputs x K = (n2, re / im); values doubled (P = 4|X|^2)
    unsigned char b2s[2][32 * 128];              // special: N = 32 outputs x K = n2 (first 64 B of each row)
    float4 tw[5 * 32];                           // [jq][TMEM lane]: (cos a, cos b, sin a, sin b) of W_2048^(n2 k1), k1 = 8 jq, 8 jq + 1;
                                                 // [4][lane]: (cos, cos, sin, sin) of W_2048^(2 n2), the rotation to the next k1 pair
    float win[32 * 36];                          // [lane][4 j + e] = Hann(128 j + 4 lane + e), j < 8; rows padded to 36
};
static_assert(sizeof(TcConst) % 16 == 0, "copied as uint4");

struct MelTab {                                  // the mel-walk part of LogmelTables; coef / gslot TRANSPOSED to [term][band]
    float2 coef[kMel * kMaxTerms];               // so that the lanes (= bands) of the closing loop read consecutive words
    unsigned long long lanemask[32];
    unsigned char gslot[kMel * kMaxTerms];
    unsigned char lanebase[32];
    int terms_round1, terms_round2;
    int pad_[2];
};

// shared-memory map (bytes from the 1024-aligned base).  The special-row atoms come first: their MMA reads 128 rows,
// i.e. 16 KB from the atom's base, of which only rows 0..3 are meaningful -- the rest must merely be mapped.
constexpr int kOffSpec = 0;
constexpr int kOffTiles = kOffSpec + kTeams * kSpecBytes;           // 8192
constexpr int kOffConst = kOffTiles + kTeams * kTileBytes;          // 139264 (1024-aligned)
constexpr int kOffPart = kOffConst + (int)sizeof(TcConst);
constexpr int kOffMel = kOffPart + kTeams * 4 * kMaxSlots * 8;
constexpr int kOffBars = kOffMel + (int)sizeof(MelTab);
constexpr int kTcSmem = kOffBars + kTeams * 8 + 16 + 1024;
static_assert(kOffTiles % 1024 == 0 && kOffConst % 1024 == 0, "swizzle atoms are 1024-byte aligned");
static_assert(kOffBars % 8 == 0 && kOffMel % 8 == 0 && kOffPart % 16 == 0, "alignment");
static_assert(4 * kPStride * 4 <= kTileBytes, "the four power spectra of a tile live in its operand buffer");
static_assert(kTcSmem <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t sw128(int r, int c) {          // byte offset of 16-byte chunk c of row r
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// (a, b) -> fp16 pair hi (low half = a) and the fp16 pair of the remainders a - hi_a, b - hi_b:
// F2FP, two mixed-precision FHFMA (fp32 = fp16 * -1 + fp32), F2FP
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    float la, lb;
    asm("{\n\t.reg .b16 l, h, m;\n\t"
        "cvt.rn.f16x2.f32 %0, %4, %3;\n\t"
        "mov.b32 {l, h}, %0;\n\t"
        "mov.b16 m, 0xBC00;\n\t"
        "fma.rn.f32.f16 %1, l, m, %3;\n\t"
        "fma.rn.f32.f16 %2, h, m, %4;\n\t}"
        : "=&r"(hi), "=f"(la), "=f"(lb)
        : "f"(a), "f"(b));
    asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(lo) : "f"(la), "f"(lb));
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// mbarrier wait that SLEEPS in hardware (suspend-time hint) instead of polling: a polling role eats the issue slots of the
// roles it is waiting for (first build: 700 of 1840 warp instructions per frame were wait loops).  Bounded: a pipeline
// bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < 200000; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
    }
    __trap();
}
// shared-memory accesses by 32-bit shared address (generic pointers cost 64-bit address arithmetic and LD / ST)
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float2 lds64(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds128c(uint32_t a) {      // constants: may be hoisted / reordered
    float4 v; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float2 lds64c(uint32_t a) { float2 v; asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8c(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void stsf32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ bool elect_one() {                 // one lane of a converged warp
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((unsigned short)a) : "memory");
}

// four consecutive samples (16-byte / 8-byte aligned)
__device__ __forceinline__ float4 ld_quad(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_quad(const short* p) {
    const short4 v = __ldg(reinterpret_cast<const short4*>(p));
    constexpr float k = 1.0f / 32768.0f;
    return make_float4((float)v.x * k, (float)v.y * k, (float)v.z * k, (float)v.w * k);
}

template <typename T>
__global__ void __launch_bounds__(kTcThreads, 1)
logmel_tc_kernel(const T* __restrict__ pcm, float* __restrict__ out, int n_ch, long S, unsigned n_frames,
                 unsigned total_frames, int n_tiles, int pad_mode, const TcConst* __restrict__ gconst,
                 const LogmelTables* __restrict__ gtab) {
    extern __shared__ unsigned char lmtc_raw[];
    // 1024-byte alignment by pointer arithmetic ON the shared array (an integer round trip would turn every access below
    // into a generic LD / ST with 64-bit address arithmetic instead of LDS / STS with immediate offsets)
    unsigned char* smem = lmtc_raw + ((1024u - (smem_u32(lmtc_raw) & 1023u)) & 1023u);
    TcConst& cst = *reinterpret_cast<TcConst*>(smem + kOffConst);
    MelTab& mel = *reinterpret_cast<MelTab*>(smem + kOffMel);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);              // one per team: "my MMAs have retired"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTeams);
    const int tid = threadIdx.x, lane = tid & 31;
    // the shuffle tells the compiler that the warp index is warp-uniform: team / tile addresses / MMA descriptors then
    // live in uniform registers and every tcgen05.mma is ONE instruction (as thread registers each one was wrapped in
    // a 14-instruction ELECT / R2UR / branch loop on the team's critical path)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int team = warp >> 2, fw = warp & 3;                                  // fw: frame of the tile = TMEM sub-partition

    if (tid == 0) {
        for (int i = 0; i < kTeams; ++i) mbar_init(bars + i, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {
        const uint4* src = reinterpret_cast<const uint4*>(gconst);
        uint4* dst = reinterpret_cast<uint4*>(&cst);
        for (int i = tid; i < (int)(sizeof(TcConst) / 16); i += kTcThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < kMel * kMaxTerms; i += kTcThreads) {
            const int bnd = i / kMaxTerms, term = i - bnd * kMaxTerms;
            mel.coef[term * kMel + bnd] = gtab->coef[i];
            mel.gslot[term * kMel + bnd] = gtab->gslot[i];
        }
        if (tid < 32) { mel.lanemask[tid] = gtab->lanemask[tid]; mel.lanebase[tid] = gtab->lanebase[tid]; }
        if (tid == 0) { mel.terms_round1 = gtab->terms_round1; mel.terms_round2 = gtab->terms_round2; }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + team * kTmemCols;
    const int t0 = (int)((long)blockIdx.x * n_tiles / gridDim.x), t1 = (int)((long)(blockIdx.x + 1) * n_tiles / gridDim.x);
    const int n_local = t1 - t0;
    uint64_t* bar = bars + team;
    uint32_t phase = 0;
    const bool issuer_warp = fw == 0;

    // ---- per-thread constants of the three phases
    unsigned char* const tilep = smem + kOffTiles + team * kTileBytes;          // A1, then A2, then P of the team's tile
    unsigned char* const specp = smem + kOffSpec + team * kSpecBytes;
    const uint32_t tile = smem_u32(tilep), spec = smem_u32(specp);
    // producer: rows 8 e + (lane & 7) of the frame, 16-byte chunks 2 (lane >> 3) + h
    const int m = lane & 7, g = lane >> 3;
    const float4* const wrow = reinterpret_cast<const float4*>(cst.win + lane * 36);
    unsigned char* const a1row = tilep + (4 * fw) * 1024 + m * 128;
    // stage-1 epilogue: TMEM lane = row position of (frame, n2)
    const int n2 = 4 * m + g;
    unsigned short* const spec_row = reinterpret_cast<unsigned short*>(specp + fw * 128 + (((n2 >> 3) ^ fw) << 4) + (n2 & 7) * 2);
    const float4* const twp = cst.tw + lane;
    const float2 rot_c = make_float2(twp[4 * 32].x, twp[4 * 32].y), rot_s = make_float2(twp[4 * 32].z, twp[4 * 32].w);
    unsigned char* const a2rows = tilep + (4 * fw) * 1024 + g * 4;              // + jq * 1024 + q * 128 + ((m ^ q) << 4)
    const uint32_t ta = tmem + ((uint32_t)(32 * fw) << 16);
    // stage-2 epilogue + mel
    float* const P = reinterpret_cast<float*>(tilep) + fw * kPStride;
    float* const Pfwd = P + lane;                                               // bin k1 + 64 k2
    float* const Pmir = P + 2048 - lane;                                        // mirror bin 2048 - k
    float2* const part = reinterpret_cast<float2*>(smem + kOffPart) + warp * kMaxSlots;
    const unsigned long long msk = mel.lanemask[lane];
    const uint32_t slot0 = smem_u32(part + mel.lanebase[lane]);
    const int terms1 = mel.terms_round1, terms2 = mel.terms_round2;
    // MMA operands (the issuing thread)
    constexpr uint32_t idesc64 = idesc_f16(128, 64, 0, 0), idesc32 = idesc_f16(128, 32, 0, 0);
    const uint64_t dzero = smem_desc_sw128(0, 16, 1024);
    const uint64_t dAh = dzero | (uint64_t)((tile & 0x3FFFFu) >> 4), dAl = dzero | (uint64_t)(((tile + kPlane) & 0x3FFFFu) >> 4);
    const uint64_t dSh = dzero | (uint64_t)((spec & 0x3FFFFu) >> 4), dSl = dzero | (uint64_t)(((spec + 1024) & 0x3FFFFu) >> 4);
    const uint64_t dB1h = dzero | (uint64_t)((smem_u32(cst.b1[0]) & 0x3FFFFu) >> 4), dB1l = dzero | (uint64_t)((smem_u32(cst.b1[1]) & 0x3FFFFu) >> 4);
    const uint64_t dB2h = dzero | (uint64_t)((smem_u32(cst.b2[0]) & 0x3FFFFu) >> 4), dB2l = dzero | (uint64_t)((smem_u32(cst.b2[1]) & 0x3FFFFu) >> 4);
    const uint64_t dBsh = dzero | (uint64_t)((smem_u32(cst.b2s[0]) & 0x3FFFFu) >> 4), dBsl = dzero | (uint64_t)((smem_u32(cst.b2s[1]) & 0x3FFFFu) >> 4);

    // ---- phase 1 of tile `ti` of this team: PCM -> Hann -> per-frame scale, left in v[] (64 registers).  It runs one
    //      tile AHEAD, between the issue of a tile's stage-2 MMAs and the wait for them, so the tensor-core time and the
    //      DRAM latency of the loads hide behind each other.
    float4 v[16];
    bool n_live = false;
    unsigned n_cc = 0, n_frame = 0;
    float n_unscale = 0.0f;
    auto load_frame = [&](int ti) {
        const unsigned q0 = 4u * (unsigned)(t0 + ti) + (unsigned)fw;
        n_live = q0 < total_frames;
        const unsigned q = n_live ? q0 : total_frames - 1;        // a dead row of the last tile recomputes the last frame
        const unsigned cc = q / n_frames, frame = q - cc * n_frames;
        n_cc = cc; n_frame = frame;
        // =========================================== phase 1: PCM -> Hann -> per-frame scale -> fp16 hi / lo A1 rows
        {   // the team's next tile: pull its samples into L2 now (no registers), ~15k cycles before they are loaded
            const unsigned qn = q0 + 4u * kTeams;
            if (qn < total_frames) {
                const unsigned ccn = qn / n_frames, fn = qn - ccn * n_frames;
                const long sn = ((long)fn - 1) * kHop;
                if (sn >= 0 && sn + kNfft <= S) {
                    const T* xn = pcm + (long)ccn * S + sn + (128 / sizeof(T)) * lane;          // one 128-byte line per lane
#pragma unroll
                    for (int j = 0; j < (int)(kNfft * sizeof(T) / 4096); ++j)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + j * (4096 / sizeof(T))));
                }
            }
        }
        {
            const T* __restrict__ xb = pcm + (long)cc * S;
            const long start = ((long)frame - 1) * kHop;
            if (start >= 0 && start + kNfft <= S) {
                const T* xs = xb + start + 4 * lane;
                if ((reinterpret_cast<uintptr_t>(xs) & (4 * sizeof(T) - 1)) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = ld_quad(xs + 128 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = make_float4(ld_sample(xs + 128 * j), ld_sample(xs + 128 * j + 1), ld_sample(xs + 128 * j + 2),
                                           ld_sample(xs + 128 * j + 3));
                }
            } else {                                               // first / last frames of a clip: the padding rule
#pragma unroll 1
                for (int j = 0; j < 16; ++j) {
                    const long i0 = start + 128 * j + 4 * lane;
                    const float4 val = make_float4(padded_sample(xb, S, i0, pad_mode), padded_sample(xb, S, i0 + 1, pad_mode),
                                                   padded_sample(xb, S, i0 + 2, pad_mode), padded_sample(xb, S, i0 + 3, pad_mode));
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj)                 // v[] stays in registers: no dynamic index
                        if (jj == j) v[jj] = val;
                }
            }
        }
        // per-frame power-of-two scale: max |x| of the frame lands in [256, 512), so |Y| < 2^15 and the lo planes stay
        // clear of fp16's subnormal range
        float amax = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) amax = fmaxf(fmaxf(amax, fabsf(v[j].x)), fmaxf(fmaxf(fabsf(v[j].y), fabsf(v[j].z)), fabsf(v[j].w)));
        unsigned E = __reduce_max_sync(0xffffffffu, __float_as_uint(amax)) >> 23;
        E = E < 8u ? 8u : E;
        const float s = __uint_as_float((262u - E) << 23);
        n_unscale = -1.3862943611198906f * (float)(135 - (int)E);             // -2 ln2 * exponent, added after the log
        // Hann: w for the first half, 1 - w for the second (periodic window: w[n + 1024] = 1 - w[n])
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 w = wrow[j];
            const float2 wa = __fmul2_rn(make_float2(w.x, w.y), make_float2(s, s)), wb = __fmul2_rn(make_float2(w.z, w.w), make_float2(s, s));
            const float2 ua = __ffma2_rn(make_float2(-w.x, -w.y), make_float2(s, s), make_float2(s, s));
            const float2 ub = __ffma2_rn(make_float2(-w.z, -w.w), make_float2(s, s), make_float2(s, s));
            const float2 p0 = __fmul2_rn(make_float2(v[j].x, v[j].y), wa), p1 = __fmul2_rn(make_float2(v[j].z, v[j].w), wb);
            const float2 r0 = __fmul2_rn(make_float2(v[j + 8].x, v[j + 8].y), ua), r1 = __fmul2_rn(make_float2(v[j + 8].z, v[j + 8].w), ub);
            v[j] = make_float4(p0.x, p0.y, p1.x, p1.y);
            v[j + 8] = make_float4(r0.x, r0.y, r1.x, r1.y);
        }
    };
    if (team < n_local) load_frame(team);

    for (int i = team; i < n_local; i += kTeams) {
        const bool live = n_live;
        const unsigned cc = n_cc, frame = n_frame;
        const float unscale = n_unscale;
        named_bar(1 + team, 128);                                  // the team's previous mel walk (it reads this buffer) is over
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 a4 = v[8 * h + 2 * jj], b4 = v[8 * h + 2 * jj + 1];
                    const float a = e == 0 ? a4.x : e == 1 ? a4.y : e == 2 ? a4.z : a4.w;
                    const float bb = e == 0 ? b4.x : e == 1 ? b4.y : e == 2 ? b4.z : b4.w;
                    split2(a, bb, hi[jj], lo[jj]);
                }
                unsigned char* addr = a1row + e * 1024 + (((2 * g + h) ^ m) << 4);
                *reinterpret_cast<uint4*>(addr) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(addr + kPlane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_proxy_async();
        named_bar(1 + team, 128);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mma_bf16(tmem, dAh + 2 * k, dB1h + 2 * k, idesc64, k != 0);
                mma_bf16(tmem, dAh + 2 * k, dB1l + 2 * k, idesc64, 1);
                mma_bf16(tmem, dAl + 2 * k, dB1h + 2 * k, idesc64, 1);
            }
            mma_commit(bar);
        }
        if (fw == 0) wait_bar(bar, phase);                         // one warp polls; the others sleep at the barrier
        phase ^= 1;
        named_bar(1 + team, 128);
        tc_fence_after();
        // =========================================== phase 2: TMEM -> twiddle -> fp16 hi / lo A2 rows, in place over A1
        {
            float y32 = 0.0f;
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                float re[8], im[8];
                tmem_ld8_nowait(ta + 8 * jq, re);
                tmem_ld8_nowait(ta + 32 + 8 * jq, im);
                tmem_ld_wait();
                if (jq == 0) { y32 = im[0]; im[0] = 0.0f; }         // the slot of Im Y[0] (= 0) carries Y[32]
                // twiddles of k1 = 8 jq, 8 jq + 1 from the table; the next three pairs by rotating with W^(2 n2) (four
                // packed FMA-pipe instructions instead of a 4-wavefront LDS.128: the shared-memory pipe is the busy one)
                const float4 t0 = twp[jq * 32];
                float2 c2 = make_float2(t0.x, t0.y), s2 = make_float2(t0.z, t0.w);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (p > 0) {
                        const float2 nc = __ffma2_rn(make_float2(-s2.x, -s2.y), rot_s, __fmul2_rn(c2, rot_c));
                        const float2 ns = __ffma2_rn(c2, rot_s, __fmul2_rn(s2, rot_c));
                        c2 = nc; s2 = ns;
                    }
                    const float2 r2 = make_float2(re[2 * p], re[2 * p + 1]), i2 = make_float2(im[2 * p], im[2 * p + 1]);
                    const float2 zr = __ffma2_rn(i2, s2, __fmul2_rn(r2, c2));                       // (re + i im)(c - i s)
                    const float2 zi = __ffma2_rn(make_float2(-r2.x, -r2.y), s2, __fmul2_rn(i2, c2));
                    uint32_t hi, lo;
                    split2(zr.x, zi.x, hi, lo);
                    unsigned char* addr = a2rows + jq * 1024 + (2 * p) * 128 + ((m ^ (2 * p)) << 4);
                    *reinterpret_cast<uint32_t*>(addr) = hi;
                    *reinterpret_cast<uint32_t*>(addr + kPlane) = lo;
                    split2(zr.y, zi.y, hi, lo);
                    addr = a2rows + jq * 1024 + (2 * p + 1) * 128 + ((m ^ (2 * p + 1)) << 4);
                    *reinterpret_cast<uint32_t*>(addr) = hi;
                    *reinterpret_cast<uint32_t*>(addr + kPlane) = lo;
                }
            }
            uint32_t hi, lo;
            split2(y32, 0.0f, hi, lo);
            spec_row[0] = (unsigned short)hi;
            spec_row[512] = (unsigned short)lo;
        }
        tc_fence_before();
        fence_proxy_async();
        named_bar(1 + team, 128);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mma_bf16(tmem, dAh + 2 * k, dB2h + 2 * k, idesc64, k != 0);
                mma_bf16(tmem, dAh + 2 * k, dB2l + 2 * k, idesc64, 1);
                mma_bf16(tmem, dAl + 2 * k, dB2h + 2 * k, idesc64, 1);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                mma_bf16(tmem + 64, dSh + 2 * k, dBsh + 2 * k, idesc32, k != 0);
                mma_bf16(tmem + 64, dSh + 2 * k, dBsl + 2 * k, idesc32, 1);
                mma_bf16(tmem + 64, dSl + 2 * k, dBsh + 2 * k, idesc32, 1);
            }
            mma_commit(bar);
        }
        if (i + kTeams < n_local) load_frame(i + kTeams);          // next tile's phase 1, behind this tile's stage-2 MMAs
        if (fw == 0) wait_bar(bar, phase);
        phase ^= 1;
        named_bar(1 + team, 128);
        tc_fence_after();
        // =========================================== phase 3: TMEM -> |X|^2 -> P of the frame (over the dead operands)
#pragma unroll
        for (int jq = 0; jq < 4; ++jq) {
            float re[8], im[8];
            tmem_ld8_nowait(ta + 8 * jq, re);
            tmem_ld8_nowait(ta + 32 + 8 * jq, im);
            tmem_ld_wait();
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                const float2 r2 = make_float2(re[2 * e2], re[2 * e2 + 1]), i2 = make_float2(im[2 * e2], im[2 * e2 + 1]);
                const float2 pw2 = __ffma2_rn(r2, r2, __fmul2_rn(i2, i2));       // 4 |X|^2 (B2 is doubled)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int k2 = 8 * jq + 2 * e2 + u;
                    const float pw = u == 0 ? pw2.x : pw2.y;
                    // bin k1 + 64 k2, or its mirror 2048 - k for k2 >= 16 (row k1 = 0 is a real sequence: its mirror
                    // outputs k2 > 16 rewrite bins 64 .. 960 with the conjugate's -- equal -- power; no predicate)
                    if (k2 < 16) Pfwd[64 * k2] = pw;
                    else Pmir[-64 * k2] = pw;
                }
            }
        }
        if (lane < 31) P[1025 + lane] = 0.0f;                       // the walk reads 33 x 32 bins
        if (fw == 0) {                                             // special rows: TMEM lanes 0..3 = the tile's frames
            float sp[32];
            tmem_ld32(tmem + 64, sp);
            if (lane < 4) {
                float* Ps = reinterpret_cast<float*>(tilep) + lane * kPStride + 32;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) Ps[64 * k2] = fmaf(sp[k2], sp[k2], sp[16 + k2] * sp[16 + k2]);
            }
        }
        tc_fence_before();
        named_bar(1 + team, 128);                                  // every bin of the four frames is in place
        if (!live) continue;
        // ---- mel projection: the walk of logmel.cu (per-lane (sum P, sum i P) per band-edge segment, fixed order)
        {
            const float* Pl = P + lane * kBinStride;
            const unsigned mlo = (unsigned)msk, mhi = (unsigned)(msk >> 32);
            unsigned slot = slot0;
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int ii = 0; ii < kBinStride; ++ii) {
                const float p = Pl[ii];
                if (ii > 0) {
                    asm volatile(
                        "{\n\t.reg .pred q;\n\t"
                        "setp.ne.u32 q, %3, 0;\n\t"
                        "@q st.shared.v2.f32 [%0], {%1, %2};\n\t"
                        "@q add.u32 %0, %0, 8;\n\t"
                        "@q mov.f32 %1, 0f00000000;\n\t"
                        "@q mov.f32 %2, 0f00000000;\n\t}"
                        : "+r"(slot), "+f"(s0), "+f"(s1)
                        : "r"((ii < 32 ? mlo : mhi) & (1u << (ii & 31)))
                        : "memory");
                }
                s0 += p;
                s1 = fmaf((float)ii, p, s1);
            }
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(s0), "f"(s1) : "memory");
        }
        __syncwarp();
        {
            const unsigned clip = cc / (unsigned)n_ch, ch = cc - clip * (unsigned)n_ch;
            float* o = out + (((long)clip * n_frames + frame) * n_ch + ch) * kMel;
            {
                const int bnd = kMel - kBandsRound1 + lane;
                float acc = 0.0f;
                for (int ii = 0; ii < terms1; ++ii) {
                    const float2 c = mel.coef[ii * kMel + bnd];
                    const float2 sv = part[mel.gslot[ii * kMel + bnd]];
                    acc = fmaf(c.x, sv.x, acc);
                    acc = fmaf(c.y, sv.y, acc);
                }
                o[bnd] = logf(acc) + unscale;
            }
            if (lane < kMel - kBandsRound1) {
                const int bnd = lane;
                float acc = 0.0f;
                for (int ii = 0; ii < terms2; ++ii) {
                    const float2 c = mel.coef[ii * kMel + bnd];
                    const float2 sv = part[mel.gslot[ii * kMel + bnd]];
                    acc = fmaf(c.x, sv.x, acc);
                    acc = fmaf(c.y, sv.y, acc);
                }
                o[bnd] = logf(acc) + unscale;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(*tmem_slot, 512);
    }
}

// ------------------------------------------------------------------------------ host: constant operands
void put_f16(unsigned char* plane_hi, unsigned char* plane_lo, int n, int k, double v) {
    const __half h = __float2half_rn((float)v);
    const __half l = __float2half_rn((float)(v - (double)__half2float(h)));
    const size_t off = (size_t)((n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
    std::memcpy(plane_hi + off, &h, 2);
    std::memcpy(plane_lo + off, &l, 2);
}

void build_const(TcConst& c) {
    std::memset(&c, 0, sizeof(c));
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < 64; ++n)                                   // stage 1, output column n
        for (int kp = 0; kp < 64; ++kp) {
            const int n1 = (kp >> 4) + 4 * (kp & 15);              // K position -> n1 (the producers' float4 order)
            double v;
            if (n < 32) v = std::cos(two_pi * (double)((n1 * n) % 64) / 64.0);
            else if (n == 32) v = (n1 & 1) ? -1.0 : 1.0;
            else v = -std::sin(two_pi * (double)((n1 * (n - 32)) % 64) / 64.0);
            put_f16(c.b1[0], c.b1[1], n, kp, v);
        }
    for (int n = 0; n < 64; ++n)                                   // stage 2, doubled: P comes out as 4 |X|^2
        for (int k = 0; k < 64; ++k) {
            const int n2 = k >> 1, r = k & 1, k2 = n & 31;
            const double th = two_pi * (double)((n2 * k2) % 32) / 32.0;
            const double v = n < 32 ? (r == 0 ? std::cos(th) : std::sin(th)) : (r == 0 ? -std::sin(th) : std::cos(th));
            put_f16(c.b2[0], c.b2[1], n, k, 2.0 * v);
        }
    for (int n = 0; n < 32; ++n)                                   // special rows: X[32 + 64 k2] = sum_n2 Y32[n2] W_64^(n2 (2 k2 + 1))
        for (int n2 = 0; n2 < 32; ++n2) {
            const int k2 = n & 15;
            const double ph = two_pi * (double)((n2 * (2 * k2 + 1)) % 64) / 64.0;
            put_f16(c.b2s[0], c.b2s[1], n, n2, 2.0 * (n < 16 ? std::cos(ph) : -std::sin(ph)));
        }
    for (int l = 0; l < 32; ++l) {
        const int n2 = 4 * (l & 7) + (l >> 3);
        for (int jq = 0; jq < 4; ++jq) {
            const double a = two_pi * (double)(n2 * (8 * jq)) / kNfft, b = two_pi * (double)(n2 * (8 * jq + 1)) / kNfft;
            c.tw[jq * 32 + l] = make_float4((float)std::cos(a), (float)std::cos(b), (float)std::sin(a), (float)std::sin(b));
        }
        const double r = two_pi * (double)(2 * n2) / kNfft;
        c.tw[4 * 32 + l] = make_float4((float)std::cos(r), (float)std::cos(r), (float)std::sin(r), (float)std::sin(r));
    }
    for (int l = 0; l < 32; ++l)
        for (int j = 0; j < 8; ++j)
            for (int e = 0; e < 4; ++e) {
                const int n = 128 * j + 4 * l + e;
                c.win[l * 36 + 4 * j + e] = (float)(0.5 - 0.5 * std::cos(two_pi * n / kNfft));
            }
}

std::mutex g_tc_mu;
std::map<int, TcConst*> g_tc_const;              // device -> device copy

int get_const(cudaStream_t stream, const TcConst** out) {
    int dev = 0;
    SED_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tc_mu);
    auto it = g_tc_const.find(dev);
    if (it == g_tc_const.end()) {
        static TcConst host;                      // guarded by g_tc_mu
        build_const(host);
        TcConst* d = nullptr;
        SED_CUDA_OK(cudaMalloc(&d, sizeof(TcConst)));
        SED_CUDA_OK(cudaMemcpyAsync(d, &host, sizeof(TcConst), cudaMemcpyHostToDevice, stream));
        SED_CUDA_OK(cudaStreamSynchronize(stream));   // one-off
        it = g_tc_const.emplace(dev, d).first;
    }
    *out = it->second;
    return SEDB200_OK;
}

}  // namespace

template <typename T>
int logmel_tc_launch(const T* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode, float* out_dev,
                     cudaStream_t st) {
    const LogmelTables* tab = nullptr;
    int rc = logmel_get_tables(sr, st, &tab);
    if (rc) return rc;
    const TcConst* cst = nullptr;
    rc = get_const(st, &cst);
    if (rc) return rc;
    const long nfr = 1 + n_samples / kHop;
    const long total = (long)n_clips * n_ch * nfr;
    SED_REQUIRE(nfr < (1L << 31) && total < (1L << 31), SEDB200_ESHAPE, "logmel (tensor-core kernel): %ld frames in one launch", total);
    const int n_tiles = (int)((total + 3) / 4);
    const int grid = std::min(n_tiles, sm_count());
    rc = ensure_dyn_smem((const void*)logmel_tc_kernel<T>, kTcSmem);
    if (rc) return rc;
    logmel_tc_kernel<T><<<grid, kTcThreads, kTcSmem, st>>>(pcm_dev, out_dev, n_ch, n_samples, (unsigned)nfr, (unsigned)total,
                                                           n_tiles, pad_mode, cst, tab);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

template int logmel_tc_launch<float>(const float*, int, int, long, int, int, float*, cudaStream_t);
template int logmel_tc_launch<short>(const short*, int, int, long, int, int, float*, cudaStream_t);

}  // namespace sedb200
