// tc_conv.cu -- 3x3 convolutions as implicit GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA, fp32-grade accuracy through a 3-term bf16 split.
//
// Reference call sites: nn.Conv2d(.., 3, padding=1) in crnn_lightning.py:47 / sed.py:88 and their
// autograd backward (dgrad / wgrad).
//
// Why split precision: the parity gate (probabilities within 1e-3 after one Adam step) is not met by
// single-pass bf16 (2-3e-3) nor TF32 (1.2e-3) operands -- Adam's first step moves every weight by
// lr*sign(g), so operand rounding in the gradients is amplified (DESIGN.md "precision").  Every fp32
// operand x is therefore carried as two bf16 planes  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi),
// 16 mantissa bits together) and each product runs as three MMAs into the same fp32 TMEM accumulator:
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi            (the dropped lo*lo term is ~2^-18 relative)
//
// Forward / dgrad kernel (K-major operands):
//   M = 128 output pixels of one image (Ht = 128/W rows x W columns), N = 128 output channels,
//   K = 9 taps x Cin, walked as K-blocks of one tap x 64 channels.  For a tap (r,s) the A tile is the
//   SAME NHWC box shifted by (r-1, s-1): a 4-D TMA load {64 c, W, Ht, 1} at coordinates
//   (c0, s-1, h0+r-1, b); out-of-image rows/columns are zero-filled by the TMA unit, which is exactly
//   the conv padding.  No im2col buffer ever exists.
//   Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//   warps 4-7 = epilogue (tcgen05.ld -> +bias -> fp32 NHWC store).  Two TMEM accumulator buffers so the
//   epilogue of tile i overlaps the MMAs of tile i+1; 3-stage smem ring (64 KB per stage).
//   dgrad is the same kernel on dY with the weights flipped and transposed beforehand.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "tc_umma.cuh"
#include "tc_conv.cuh"
#include "crnn_block.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace sedb200 {

// ---------------------------------------------------------------------------------- tensor-map encode
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            SED_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
            SED_REQUIRE(p && q == cudaDriverEntryPointSuccess, SEDB200_ECUDA, "cuTensorMapEncodeTiled not available");
            fn = reinterpret_cast<Fn>(p);
        }
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                    strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SED_REQUIRE(r == CUDA_SUCCESS, SEDB200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SEDB200_OK;
}

namespace {
using namespace umma;

constexpr int kStages = 3;                                       // wgrad_tc_kernel, 3-term
constexpr int kMaxStages = 6;                                    // barrier slots
constexpr int kTileM = 128, kTileN = 128, kBlockK = 64;
constexpr int kTileBytes = kTileM * kBlockK * 2;                 // 16 KB: one 16-bit operand tile
constexpr int kStageBytes = 4 * kTileBytes;                      // wgrad_tc_kernel: A_hi, A_lo, B_hi, B_lo
constexpr int kEpiPitch = 36;                                    // floats per staged row: LDS/STS.128 conflict-free
constexpr int kEpiTile = 32 * kEpiPitch * 4;                     // one 32 x 32 staging tile per epilogue warp
constexpr int kThreads = 256;                                    // wgrad_tc_kernel
constexpr int kConvThreads = 384;                                // conv_tc_kernel: warps 8-11 = second epilogue set
constexpr uint32_t kTmemCols = 512;                              // conv_tc_kernel: 2 buffers x 2 tiles x 128 columns

// conv_tc_kernel modes.  Every tensor-core kernel here turned out to be bound by the bytes an SM can ingest from L2
// (~50 B/clk/SM measured: profiles/README.md), so the knobs are about bytes per MMA:
//   terms  3: A_hi*B_hi + A_hi*B_lo + A_lo*B_hi (fp32-grade, forward) / 1: hi planes only (gradient contractions)
//   tpi    tiles per work item: 2 = two M tiles (256 output pixels) share every weight k-block -- 24 instead of 32 KB
//          per 128x128x64 MMA block -- when there are enough tiles to keep the 148 CTAs balanced; 1 otherwise
// stage = tpi * planes A tiles + planes B tiles of 16 KB, planes = 2 (hi, lo) or 1.
struct ConvMode { int planes, tpi, stage_bytes, n_stages, n_epi_warps, smem_bytes; };
inline ConvMode conv_mode(int terms, int tpi) {
    ConvMode m;
    m.planes = terms >= 2 ? 2 : 1;
    m.tpi = tpi;
    m.stage_bytes = (tpi + 1) * m.planes * kTileBytes;           // 32 / 48 / 64 / 96 KB
    // 1-term: a third of the MMA work per tile -> EIGHT epilogue warps (two per TMEM sub-partition, half of the
    // columns each); the 4-warp epilogue was the bound there (ncu: tensor pipe 40 %, issue 16 %)
    m.n_epi_warps = terms == 1 ? 8 : 4;
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + 4096 /*BN-stat staging*/ + m.n_epi_warps * kEpiTile;
    m.n_stages = std::min(kMaxStages, (227 * 1024 - fixed) / m.stage_bytes);
    m.smem_bytes = m.n_stages * m.stage_bytes + fixed;
    return m;
}

// Halo mode: two thirds of the A bytes a launch pulls from L2 are the same pixels again: the boxes of the taps (r, s),
// r = 0..2, are ONE box of Ht + 2 image rows seen from
// three row offsets.  A row offset is W x 128 bytes, a whole number of 1024-byte swizzle atoms when W % 8 == 0, so the
// three taps are three descriptor start addresses into the same box (no base-offset field needed).  Two rings instead
// of one: A slots hold {planes} boxes of (tpi*Ht + 2) x W rows for one (s, 64-channel chunk), B slots hold one tap's
// weight tile(s); the K walk is (s, chunk, r).  With two time-adjacent M tiles per item (single-pass launches) a
// 128x128x64 MMA block costs 6 + 8 = 14 KB of shared-memory fill instead of 24 (forward, tpi 1: 45 instead of 64).
struct HaloMode { int n_a, n_b, a_plane_bytes, a_slot_bytes, b_slot_bytes, ring_bytes, n_epi_warps, smem_bytes; };
inline HaloMode conv_halo_mode(int terms, int tpi, int W) {
    HaloMode m;
    const int planes = terms >= 2 ? 2 : 1;
    m.a_plane_bytes = (tpi * kTileM + 2 * W) * 128;
    m.a_slot_bytes = planes * m.a_plane_bytes;
    m.b_slot_bytes = planes * kTileBytes;
    m.n_epi_warps = terms == 1 ? 8 : 4;
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + 4096 /*BN-stat staging*/ + m.n_epi_warps * kEpiTile;
    m.n_a = 2;
    m.n_b = std::min(kMaxStages, (227 * 1024 - fixed - m.n_a * m.a_slot_bytes) / m.b_slot_bytes);
    m.ring_bytes = m.n_a * m.a_slot_bytes + std::max(m.n_b, 0) * m.b_slot_bytes;
    m.smem_bytes = m.ring_bytes + fixed;
    return m;
}

struct ConvTcParams {
    int B, H, W, lgW, Ht, tiles_per_img, n_tiles_n, m_tiles, total_items, kchunks, n_total;   // W = 2^lgW (a divisor of 128)
    int terms, tpi, planes, stage_bytes, n_stages, n_epi_warps;
    // halo mode (conv_halo_mode): one A box per (tap column, channel chunk) carries the rows of all three tap rows
    int halo, groups_per_img, n_a, a_plane_bytes, a_slot_bytes, ring_bytes;
    uint32_t idesc;          // instruction descriptor (operand formats: bf16 or fp16 planes)
    const float* out_scale;  // null, or a device scalar every accumulator is multiplied by (fp16 gradient planes)
    const float* out2_scale; // terms == 2: device scalar of the fp8 correction accumulator
    float* out;              // [B*H*W][out_ld] fp32
    const float* bias;       // [n_total] or null
    long out_ld;
    float* stats;            // null, or per-M-tile partial BatchNorm sums [m_tiles][2][n_total] (sum, sum of squares)
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const ConvTcParams p) {
    pdl_wait();
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_stages = p.n_stages, stage_bytes = p.stage_bytes, n_epi_warps = p.n_epi_warps, tpi = p.tpi, planes = p.planes;
    const int ring_bytes = p.halo ? p.ring_bytes : n_stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_bytes);
    uint64_t* full = bars;                       // [kMaxStages]  TMA -> MMA   (halo mode: the B ring)
    uint64_t* empty = bars + kMaxStages;         // [kMaxStages]  MMA -> TMA
    uint64_t* tfull = bars + 2 * kMaxStages;     // [2]           MMA -> epilogue
    uint64_t* tempty = bars + 2 * kMaxStages + 2;// [2]           epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
    uint64_t* afull = bars + 2 * kMaxStages + 6; // [2]           halo mode: the A ring
    uint64_t* aempty = bars + 2 * kMaxStages + 8;// [2]
    float* stat_s = reinterpret_cast<float*>(smem + ring_bytes + 256);                // [4 warps][2][128]
    float* epi_stage = stat_s + 1024;                                                 // [epilogue warps][32][kEpiPitch]

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, n_epi_warps); }
        for (int i = 0; i < 2; ++i) { mbar_init(afull + i, 1); mbar_init(aempty + i, 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int n_kb = 9 * p.kchunks;
    // work item -> (N tile, first M tile); an item covers M tiles mt0 .. mt0 + tpi - 1 (a tile past the end reads zeros
    // -- TMA fills boxes outside the tensor -- and is never stored)
    const int b_off = tpi * planes * kTileBytes;                 // B planes behind the A tiles of a stage

    if (warp == 0 && p.halo) {
        // ================= TMA producer, halo mode: lane = plane; per (s, chunk) one A box, then the three taps' B tiles
        int bs = 0; uint32_t bph = 0; int as = 0; uint32_t aph = 0;
        const bool mine = lane < planes;
        const CUtensorMap* tmA = lane ? &tmA_lo : &tmA_hi;
        const CUtensorMap* tmB = lane ? &tmB_lo : &tmB_hi;
        unsigned char* const b_ring = smem + p.n_a * p.a_slot_bytes;
        const int b_slot_bytes = planes * kTileBytes;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int nt = item % p.n_tiles_n, g = item / p.n_tiles_n;
            const int tb = g / p.groups_per_img, th = (g - tb * p.groups_per_img) * tpi * p.Ht;
            const int ncol = nt * kTileN;
            for (int s = 0; s < 3; ++s) {
                for (int c0 = 0; c0 < p.kchunks * kBlockK; c0 += kBlockK) {
                    if (lane == 0) {
                        mbar_wait(aempty + as, aph ^ 1);
                        mbar_expect_tx(afull + as, p.a_slot_bytes);
                    }
                    __syncwarp();
                    if (mine) tma_load_4d(smem + as * p.a_slot_bytes + lane * p.a_plane_bytes, tmA, afull + as, c0, s - 1, th - 1, tb);
                    if (++as == p.n_a) { as = 0; aph ^= 1; }
                    for (int r = 0; r < 3; ++r) {
                        if (lane == 0) {
                            mbar_wait(empty + bs, bph ^ 1);
                            mbar_expect_tx(full + bs, b_slot_bytes);
                        }
                        __syncwarp();
                        if (mine) tma_load_2d(b_ring + bs * b_slot_bytes + lane * kTileBytes, tmB, full + bs, c0, (r * 3 + s) * p.n_total + ncol);
                        if (++bs == n_stages) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && p.halo) {
        // ================= MMA issuer, halo mode: tap row r = descriptor start r * W rows into the A box =============
        const uint32_t idesc = p.idesc;
        const uint64_t dbase = smem_desc_sw128(smem_u32(smem), 16, 1024);
        const uint32_t a_slot_u = (uint32_t)p.a_slot_bytes >> 4, a_plane_u = (uint32_t)p.a_plane_bytes >> 4;
        const uint32_t b_ring_u = (uint32_t)(p.n_a * p.a_slot_bytes) >> 4, b_slot_u = (uint32_t)(planes * kTileBytes) >> 4;
        const uint32_t tile_u = kTileBytes >> 4, row_u = (uint32_t)(p.W * 128) >> 4;
        int bs = 0; uint32_t bph = 0; int as = 0; uint32_t aph = 0;
        int buf = 0; uint32_t bphase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            mbar_wait(tempty + buf, bphase ^ 1);
            tc_fence_after();
            uint32_t acc0 = 0;
            for (int sc = 0; sc < 3 * p.kchunks; ++sc) {
                mbar_wait(afull + as, aph);
                tc_fence_after();
                const uint64_t da = dbase + (uint64_t)(as * a_slot_u);
                for (int r = 0; r < 3; ++r) {
                    mbar_wait(full + bs, bph);
                    tc_fence_after();
                    const uint64_t dbh0 = dbase + (uint64_t)(b_ring_u + bs * b_slot_u), dbl0 = dbh0 + tile_u;
                    const uint64_t dar = da + (uint64_t)(r * row_u);
                    if (elect_one()) {
                    if (p.terms == 2) {
                        const uint32_t d1 = tmem_base + (buf * 2) * kTileN, d2 = d1 + kTileN;
                        const uint64_t dac = dar + a_plane_u;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) mma_bf16(d1, dar + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_f8(d2, dac + 2 * k, dbl0 + 2 * k, idesc, k ? 1u : acc0);
                    } else if (planes == 1) {
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (t < tpi) {
                                const uint32_t d = tmem_base + (buf * 2 + t) * kTileN;
                                const uint64_t dah0 = dar + (uint64_t)(t * tile_u);      // tile t: Ht rows = 16 KB further
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) mma_bf16(d, dah0 + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
                            }
                        }
                    } else {                                                             // 3-term, one tile per item
                        const uint32_t d = tmem_base + (buf * 2) * kTileN;
                        const uint64_t dal0 = dar + a_plane_u;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            mma_bf16(d, dar + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
                            mma_bf16(d, dar + 2 * k, dbl0 + 2 * k, idesc, 1);
                            mma_bf16(d, dal0 + 2 * k, dbh0 + 2 * k, idesc, 1);
                        }
                    }
                    mma_commit(empty + bs);
                    if (r == 2) mma_commit(aempty + as);
                    if (r == 2 && sc == 3 * p.kchunks - 1) mma_commit(tfull + buf);
                    }
                    __syncwarp();
                    acc0 = 1;
                    if (++bs == n_stages) { bs = 0; bph ^= 1; }
                }
                if (++as == p.n_a) { as = 0; aph ^= 1; }
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
    } else if (warp == 0) {
        // ================= TMA producer: lane i requests box i of a stage (A tiles first, then B; hi planes on even
        //                   lanes, lo planes on odd lanes in the 3-pass mode); lane 0 waits for the slot ===============
        int stage = 0; uint32_t phase = 0;
        const int n_boxes = (tpi + 1) * planes;
        const bool mine = lane < n_boxes;
        const int opnd = lane / planes, plane = lane - opnd * planes;        // opnd < tpi: A tile, == tpi: B
        const bool is_b = opnd == tpi;
        const CUtensorMap* tm = is_b ? (plane ? &tmB_lo : &tmB_hi) : (plane ? &tmA_lo : &tmA_hi);
        const int dst_off = lane * kTileBytes;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int nt = item % p.n_tiles_n, mt0 = (item / p.n_tiles_n) * tpi;
            const int mt = mt0 + (is_b ? 0 : opnd);
            const int tb = mt / p.tiles_per_img, th = (mt - tb * p.tiles_per_img) * p.Ht;   // once per item
            const int ncol = nt * kTileN;
            for (int tap = 0; tap < 9; ++tap) {
                const int r = tap / 3, s = tap - 3 * r;    // constant divisor
                for (int c0 = 0; c0 < p.kchunks * kBlockK; c0 += kBlockK) {
                    if (lane == 0) {
                        mbar_wait(empty + stage, phase ^ 1);
                        mbar_expect_tx(full + stage, stage_bytes);
                    }
                    __syncwarp();
                    if (mine) {
                        unsigned char* st = smem + stage * stage_bytes + dst_off;
                        if (is_b) tma_load_2d(st, tm, full + stage, c0, tap * p.n_total + ncol);
                        else tma_load_4d(st, tm, full + stage, c0, s - 1, th + r - 1, tb);
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-wide control flow, one elected lane issues: tc_umma.cuh elect_one) ====
        // The issuing thread is a serial instruction stream: with one MMA pass per k-step a stage holds only ~256 cycles
        // of tensor work, so descriptor construction must not cost more than that -- every descriptor is the
        // descriptor of the shared-memory base plus a (16-byte-unit) offset added to its low word.
        const uint32_t idesc = p.idesc;
        const uint64_t dbase = smem_desc_sw128(smem_u32(smem), 16, 1024);
        const uint32_t stage_u = (uint32_t)stage_bytes >> 4, tile_u = kTileBytes >> 4, b_u = (uint32_t)b_off >> 4;
        int stage = 0; uint32_t phase = 0;
        int buf = 0; uint32_t bphase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            mbar_wait(tempty + buf, bphase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                const uint64_t ds = dbase + (uint64_t)(stage * stage_u);
                const uint64_t dbh0 = ds + b_u, dbl0 = dbh0 + tile_u;
                const uint32_t acc0 = kb != 0;
                if (elect_one()) {
                if (p.terms == 2) {
                    // fp16 hi*hi into accumulator 0, the combined e4m3 correction tiles (128 fp8 along K) into accumulator 1
                    const uint32_t d1 = tmem_base + (buf * 2) * kTileN, d2 = d1 + kTileN;
                    const uint64_t dah0 = ds, dac0 = ds + tile_u;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) mma_bf16(d1, dah0 + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma_f8(d2, dac0 + 2 * k, dbl0 + 2 * k, idesc, k ? 1u : acc0);
                } else if (planes == 1) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (t < tpi) {
                            const uint32_t d = tmem_base + (buf * 2 + t) * kTileN;
                            const uint64_t dah0 = ds + (uint64_t)(t * tile_u);
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) mma_bf16(d, dah0 + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (t < tpi) {
                            const uint32_t d = tmem_base + (buf * 2 + t) * kTileN;
                            const uint64_t dah0 = ds + (uint64_t)(t * 2 * tile_u), dal0 = dah0 + tile_u;
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                mma_bf16(d, dah0 + 2 * k, dbh0 + 2 * k, idesc, k ? 1u : acc0);
                                mma_bf16(d, dah0 + 2 * k, dbl0 + 2 * k, idesc, 1);
                                mma_bf16(d, dal0 + 2 * k, dbh0 + 2 * k, idesc, 1);
                            }
                        }
                    }
                }
                mma_commit(empty + stage);                 // smem slot reusable once these MMAs retire
                if (kb == n_kb - 1) mma_commit(tfull + buf);   // accumulators complete
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
    } else if (warp >= 4 && warp < 4 + n_epi_warps) {
        // ================= epilogue (one TMEM sub-partition per warp; with 8 warps, half of the columns each) =====
        const int q = warp & 3;                            // TMEM sub-partition == warp % 4
        const int cc0 = (n_epi_warps == 8) ? ((warp - 4) >> 2) * 2 : 0, cc1 = (n_epi_warps == 8) ? cc0 + 2 : kTileN / 32;
        int buf = 0; uint32_t bphase = 0;
        const float osc = p.out_scale ? __ldg(p.out_scale) : 1.0f;
        const float osc2 = (p.terms == 2 && p.out2_scale) ? __ldg(p.out2_scale) : 0.0f;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int nt = item % p.n_tiles_n, grp = item / p.n_tiles_n;
            mbar_wait(tfull + buf, bphase);
            tc_fence_after();
            for (int t = 0; t < tpi; ++t) {
            int mt, b, h0;
            if (p.halo) {                                  // items never straddle images: (image, group of tpi tiles)
                b = grp / p.groups_per_img;
                const int ti = (grp - b * p.groups_per_img) * tpi + t;
                if (ti >= p.tiles_per_img) break;          // uniform over the epilogue warps
                mt = b * p.tiles_per_img + ti; h0 = ti * p.Ht;
            } else {
                mt = grp * tpi + t;
                if (mt >= p.m_tiles) break;                // uniform over the epilogue warps
                b = mt / p.tiles_per_img; h0 = (mt % p.tiles_per_img) * p.Ht;
            }
            // TMEM hands every lane one output ROW (pixel); each 32 x 32 chunk goes through a per-warp staging tile so
            // that one store instruction writes four full 128 B channel segments (instead of 16 B pieces of 32 pixels),
            // and the BatchNorm column sums are plain conflict-free column reads of the same tile.
            float* dst0 = p.out + (((long)b * p.H + h0) * p.W + q * 32) * p.out_ld + nt * kTileN;
            float* stg = epi_stage + (warp - 4) * (32 * kEpiPitch);
            const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
#pragma unroll 1
            for (int cc = cc0; cc < cc1; ++cc) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + t) * kTileN + cc * 32, v);
                if (p.terms == 2) {                        // + the fp8 correction accumulator (the tile's second slot)
                    float v2[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + 1) * kTileN + cc * 32, v2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(v2[j], osc2, v[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * kEpiPitch + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + nt * kTileN + cc * 32 + sub_c));
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    float4 o = *reinterpret_cast<const float4*>(stg + (r + sub_r) * kEpiPitch + sub_c);
                    o.x = fmaf(o.x, osc, bb.x); o.y = fmaf(o.y, osc, bb.y); o.z = fmaf(o.z, osc, bb.z); o.w = fmaf(o.w, osc, bb.w);
                    if ((h0 + ((q * 32 + r + sub_r) >> p.lgW)) < p.H)
                        *reinterpret_cast<float4*>(dst0 + (long)(r + sub_r) * p.out_ld + cc * 32 + sub_c) = o;
                }
                if (p.stats) {
                    // lane L: sum / sum of squares of column cc*32 + L over this warp's 32 rows, fixed order
                    const float bl = p.bias ? __ldg(p.bias + nt * kTileN + cc * 32 + lane) : 0.0f;
                    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = fmaf(stg[r * kEpiPitch + lane], osc, bl);
                        if ((h0 + ((q * 32 + r) >> p.lgW)) < p.H) { s1 += x; s2 = fmaf(x, x, s2); }
                    }
                    stat_s[(q * 2 + 0) * 128 + cc * 32 + lane] = s1;
                    stat_s[(q * 2 + 1) * 128 + cc * 32 + lane] = s2;
                }
                __syncwarp();
            }
            if (p.stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int e = (warp - 4) * 32 + lane;                      // column of the tile
                float a = 0.0f, b2 = 0.0f;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) { a += stat_s[(w4 * 2 + 0) * 128 + e]; b2 += stat_s[(w4 * 2 + 1) * 128 + e]; }
                p.stats[((long)mt * 2 + 0) * p.n_total + nt * kTileN + e] = a;
                p.stats[((long)mt * 2 + 1) * p.n_total + nt * kTileN + e] = b2;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------- operand preparation
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// fp32 [n] -> bf16 hi/lo planes
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long n4) {
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
}

// conv weight [Cout][Cin][3][3] fp32 -> B-operand planes [9][N][K] bf16 (K contiguous)
//   fwd  : N = Cout, K = Cin, plane[tap][co][ci] = w[co][ci][tap]
//   dgrad: N = Cin,  K = Cout, plane[tap][ci][co] = w[co][ci][8 - tap]      (flipped taps)
__global__ void __launch_bounds__(256)
weight_planes_kernel(const float* __restrict__ w, int Cout, int Cin, int dgrad, int fmt, __nv_bfloat16* __restrict__ hi,
                     __nv_bfloat16* __restrict__ lo, const float* __restrict__ scale2) {
    pdl_wait();
    const long n = 9L * Cout * Cin;
    const int N = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int nn = (int)((i / K) % N);
        const int tap = (int)(i / ((long)K * N));
        const int co = dgrad ? k : nn, ci = dgrad ? nn : k, t = dgrad ? 8 - tap : tap;
        const float x = __ldg(w + ((long)co * Cin + ci) * 9 + t);
        if (fmt == kPlaneF16) {                               // fp16 hi / lo (saturating), same storage
            const __half h = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f));
            reinterpret_cast<__half*>(hi)[i] = h;
            if (scale2) {
                // combined e4m3 correction plane, the OPPOSITE order of the activations' (crnn_block.cuh):
                // row [tap][n], per 64 k: 64 x e4m3(2^(b+12) (w - hi)) then 64 x e4m3(2^b w)  (activations: x unscaled,
                // residual 2^12: both products carry 2^(b+12))
                const float sb = __ldg(scale2);
                unsigned char* row = reinterpret_cast<unsigned char*>(lo) + (i / K) * 2L * K + (k >> 6) * 128 + (k & 63);
                const __nv_fp8_storage_t l8 = __nv_cvt_float_to_fp8((x - __half2float(h)) * sb * 4096.0f, __NV_SATFINITE, __NV_E4M3);
                const __nv_fp8_storage_t h8 = __nv_cvt_float_to_fp8(x * sb, __NV_SATFINITE, __NV_E4M3);
                row[0] = l8;
                row[64] = h8;
            } else {
                reinterpret_cast<__half*>(lo)[i] = __float2half_rn(fminf(fmaxf(x - __half2float(h), -65504.0f), 65504.0f));
            }
        } else {
            split_bf16(x, hi[i], lo[i]);
        }
    }
}


// ====================================================================================== wgrad
// dW[co][ci][tap] = sum_p dY[p][co] * In[p + shift(tap)][ci]
//
// One GEMM per tap with the PIXEL index as K: both operands are "MN-major" (for a fixed pixel the
// 64 channels are contiguous), which tcgen05 consumes directly from the same NHWC planes -- no
// transposes.  A K-block is 32 pixels (Hk = 32/W image rows): dY box {64 co, W, Hk, 1} and, per tap,
// the In box shifted by the tap (TMA zero-fill = padding).  A CTA owns (co-tile, ci-tile, tap row r,
// K-slice): the three taps (r, 0..2) share every dY tile and accumulate into three 128-column TMEM
// accumulators.  Per-slice partials are reduced afterwards in a fixed order (deterministic).
constexpr int kWgKp = 32;                                   // pixels per K-block
constexpr int kWgBox = kWgKp * 128;                         // 4 KB: one {64 ch x 32 px} bf16 box
constexpr int kWgABytes = 4 * kWgBox;                       // dY: 2 channel halves x (hi, lo)
constexpr int kWgBBytes = 4 * kWgBox;                       // In (one tap): 2 channel halves x (hi, lo)
constexpr int kWgStageBytes = kWgABytes + 3 * kWgBBytes;    // 64 KB
constexpr int kWgSmemBytes = kStages * kWgStageBytes + 1024 + 256;
constexpr uint32_t kWgTmemCols = 512;

struct WgradTcParams {
    int B, H, W, Hk, hblocks_per_img, total_kblocks, n_mt, n_nt, slices, terms, n_stages;
    // halo mode (single pass, W % 4 == 0): a CTA owns a tap COLUMN s; ONE In box of Hk + 2 image rows per channel half
    // serves the three tap rows as three descriptor start addresses (r * W * 128 B) -- 4 boxes and 8 + 2 * xh_bytes
    // bytes per K-block instead of 8 boxes and 32 KB; tmX_lo then carries that box shape
    int halo, xh_bytes, stage_bytes;
    uint32_t idesc;
    float* part;             // [slices][9][Cout][Cin]
    int Cout, Cin;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                const WgradTcParams p) {
    pdl_wait();
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.n_stages * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* tfull = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
    const int n_stages = p.n_stages;
    const int stage_bytes = p.stage_bytes;
    // 1-term stage: {dY hi: 2 halves} + 3 x {In hi: 2 halves}; 3-term stage: each of those followed by its lo boxes
    const int a_bytes = p.terms == 1 ? kWgABytes / 2 : kWgABytes, b_bytes = p.terms == 1 ? kWgBBytes / 2 : kWgBBytes;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmY_hi); prefetch_tmap(&tmY_lo); prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kWgTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // work item of this CTA
    int w = blockIdx.x;
    const int slice = w % p.slices; w /= p.slices;
    const int r = w % 3; w /= 3;                                   // the CTA's tap row (halo mode: its tap COLUMN)
    const int nt = w % p.n_nt, mt = w / p.n_nt;
    const int kb_per = (p.total_kblocks + p.slices - 1) / p.slices;
    const int kb0 = slice * kb_per, kb1 = min(p.total_kblocks, kb0 + kb_per);

    if (warp == 0 && p.halo) {
        // halo mode: lanes 0-1 the dY halves, lanes 2-3 the In halves (Hk + 2 rows, shifted by the CTA's tap column)
        int stage = 0; uint32_t phase = 0;
        int b = kb0 / p.hblocks_per_img, hb = kb0 - b * p.hblocks_per_img;
        const bool mine = lane < 4, is_y = lane < 2;
        const int half = lane & 1;
        const int c_first = (is_y ? mt : nt) * 128 + half * 64;
        const int dst_off = is_y ? half * kWgBox : 2 * kWgBox + half * p.xh_bytes;
        const CUtensorMap* tm = is_y ? &tmY_hi : &tmX_lo;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int h0 = hb * p.Hk;
            if (lane == 0) {
                mbar_wait(empty + stage, phase ^ 1);
                mbar_expect_tx(full + stage, stage_bytes);
            }
            __syncwarp();
            if (mine) {
                unsigned char* st = smem + stage * stage_bytes;
                if (is_y) tma_load_4d(st + dst_off, tm, full + stage, c_first, 0, h0, b);
                else tma_load_4d(st + dst_off, tm, full + stage, c_first, r - 1, h0 - 1, b);
            }
            if (++hb == p.hblocks_per_img) { hb = 0; ++b; }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && p.halo) {
        // halo mode: tap row rr = descriptor start rr * W rows into the In box; the In halves are xh_bytes apart (LBO)
        const uint32_t idesc = p.idesc;
        const uint64_t dbase_a = smem_desc_sw128(smem_u32(smem), kWgBox, 1024);
        const uint64_t dbase_b = smem_desc_sw128(smem_u32(smem), (uint32_t)p.xh_bytes, 1024);
        const uint32_t stage_u = (uint32_t)stage_bytes >> 4, a_u = (2 * kWgBox) >> 4, row_u = (uint32_t)(p.W * 128) >> 4;
        constexpr uint32_t k_u = 2048 >> 4;
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = dbase_a + (uint64_t)(stage * stage_u);
                const uint64_t db = dbase_b + (uint64_t)(stage * stage_u + a_u);
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const uint32_t d = tmem_base + rr * 128;
#pragma unroll
                    for (int k = 0; k < kWgKp / 16; ++k)
                        mma_bf16(d, da + k * k_u, db + (uint64_t)(rr * row_u) + k * k_u, idesc, (kb != kb0 || k != 0));
                }
                mma_commit(empty + stage);
                if (kb == kb1 - 1) mma_commit(tfull);
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 0) {
        // The producer is a serial instruction stream too: per K-block it has 8 (single-pass) or 16 boxes to request,
        // ~40 cycles of issue each, against 384 cycles of MMA work -- so lane i requests box i (all lanes of the warp in
        // one instruction), lane 0 alone waits for the slot and posts the byte count, and (image, row) advance without
        // divisions.
        int stage = 0; uint32_t phase = 0;
        int b = kb0 / p.hblocks_per_img, hb = kb0 - b * p.hblocks_per_img;
        // lane -> (operand, tap, half, plane): boxes 0-1 dY hi, 2-7 In hi (tap s, half); +8: the lo planes (3-term)
        const int box = lane & 7, plane = lane >> 3;
        const bool mine = lane < (p.terms == 1 ? 8 : 16);
        const bool is_y = box < 2;
        const int half = is_y ? box : (box - 2) & 1, s = is_y ? 0 : (box - 2) >> 1;
        const int c_first = (is_y ? mt : nt) * 128 + half * 64;
        const int dst_off = is_y ? (plane * 2 + half) * kWgBox : a_bytes + s * b_bytes + (plane * 2 + half) * kWgBox;
        const CUtensorMap* tm = is_y ? (plane ? &tmY_lo : &tmY_hi) : (plane ? &tmX_lo : &tmX_hi);
        for (int kb = kb0; kb < kb1; ++kb) {
            const int h0 = hb * p.Hk;
            if (lane == 0) {
                mbar_wait(empty + stage, phase ^ 1);
                mbar_expect_tx(full + stage, stage_bytes);
            }
            __syncwarp();
            if (mine) {
                unsigned char* st = smem + stage * stage_bytes;
                if (is_y) tma_load_4d(st + dst_off, tm, full + stage, c_first, 0, h0, b);
                else tma_load_4d(st + dst_off, tm, full + stage, c_first, s - 1, h0 + r - 1, b);
            }
            if (++hb == p.hblocks_per_img) { hb = 0; ++b; }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // warp-wide control flow, one elected lane issues (tc_umma.cuh: elect_one); the descriptors of a stage are the
        // descriptor of its first box plus a 16-byte-unit offset added to the low word
        const uint32_t idesc = p.idesc;                             // both operands MN-major
        const uint64_t dbase = smem_desc_sw128(smem_u32(smem), kWgBox, 1024);
        const uint32_t stage_u = (uint32_t)stage_bytes >> 4, a_u = (uint32_t)a_bytes >> 4, b_u = (uint32_t)b_bytes >> 4;
        constexpr uint32_t lo_u = (2 * kWgBox) >> 4, k_u = 2048 >> 4;
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da_hi = dbase + (uint64_t)(stage * stage_u), da_lo = da_hi + lo_u;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const uint64_t db_hi = da_hi + a_u + (uint64_t)(s * b_u), db_lo = db_hi + lo_u;
                    const uint32_t d = tmem_base + s * 128;
#pragma unroll
                    for (int k = 0; k < kWgKp / 16; ++k) {
                        mma_bf16(d, da_hi + k * k_u, db_hi + k * k_u, idesc, (kb != kb0 || k != 0));
                        if (p.terms != 1) {
                            mma_bf16(d, da_hi + k * k_u, db_lo + k * k_u, idesc, 1);
                            mma_bf16(d, da_lo + k * k_u, db_hi + k * k_u, idesc, 1);
                        }
                    }
                }
                mma_commit(empty + stage);
                if (kb == kb1 - 1) mma_commit(tfull);
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int co = mt * 128 + q * 32 + lane;
        if (kb1 > kb0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int s = 0; s < 3; ++s) {                              // accumulator s: tap (r, s), or (s, r) in halo mode
            const int tap = p.halo ? s * 3 + r : r * 3 + s;
            float* dst = p.part + (((long)slice * 9 + tap) * p.Cout + co) * p.Cin + nt * 128;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                float v[32];
                if (kb1 > kb0) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + s * 128 + cc * 32, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + cc * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kWgTmemCols);
    }
}

// dW[co][ci][tap] = sum_slice part[slice][tap][co][ci]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, int slices, int Cout, int Cin, float* __restrict__ dw,
                    const float* __restrict__ out_scale) {
    pdl_wait();
    const long n = 9L * Cout * Cin;
    const float osc = out_scale ? __ldg(out_scale) : 1.0f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        const int tap = (int)(i / ((long)Cin * Cout));
        dw[((long)co * Cin + ci) * 9 + tap] = ordered_sum<8, float>(part + i, n, slices) * osc;
    }
}

}  // namespace

// {2^b, 2^-(12 + b)}: 2^b places the largest |w| just below e4m3's 448 / 2 (the fp8 scale of a weight tensor)
__global__ void __launch_bounds__(1024)
weight_scale_kernel(const float* __restrict__ w, long n, float* __restrict__ out) {
    pdl_wait();
    __shared__ float sm[32];
    float m = 0.0f;
    const long n4 = n >> 2;                                        // conv weights: 9 * Cout * Cin floats, 16-byte aligned
    for (long i = threadIdx.x; i < n4; i += 1024) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(__ldg(w + i)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 32; ++i) m = fmaxf(m, sm[i]);
        int e = 0;
        float sb = 1.0f;
        if (m > 0.0f && isfinite(m)) {
            frexpf(m, &e);                                         // m = f * 2^e, f in [0.5, 1)
            sb = exp2f((float)max(-100, min(100, 7 - e)));         // m * sb in [64, 128)
        }
        out[0] = sb;
        out[1] = 1.0f / (sb * 4096.0f);
    }
}

size_t conv_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout) {
    const size_t act = (size_t)B * H * W * Cin * 2;      // one bf16 plane of the input
    const size_t wp = (size_t)9 * Cout * Cin * 2;
    return 2 * ((act + 1023) & ~(size_t)1023) + 2 * ((wp + 1023) & ~(size_t)1023);
}

bool conv_tc_supported(int H, int W, int Cin, int Cout) {
    return Cin % 64 == 0 && Cout % 128 == 0 && W >= 1 && W <= 128 && (128 % W) == 0 && H >= 1;
}

size_t conv_tc_weight_scratch_bytes(int Cin, int Cout) { return 2 * (((size_t)9 * Cout * Cin * 2 + 1023) & ~(size_t)1023); }

// B-operand planes of a conv weight (forward or flipped / transposed for the data gradient) into `wplanes`
// (conv_tc_weight_scratch_bytes); they stay valid until the weight changes
int conv_tc_weight_planes(const float* w, int Cin, int Cout, int dgrad, void* wplanes, cudaStream_t st, int fmt,
                          float* scale2) {
    SED_REQUIRE(!scale2 || (fmt == kPlaneF16 && !dgrad && Cin % 64 == 0), SEDB200_EINVAL,
                "conv_tc_weight_planes: the fp8 correction plane exists for the fp16 forward layout only");
    if (scale2) {
        launch_k(weight_scale_kernel, 1, 1024, 0, st, w, 9L * Cout * Cin, scale2);
        SED_POST_LAUNCH();
    }
    const size_t wp = ((size_t)9 * Cout * Cin * 2 + 1023) & ~(size_t)1023;
    __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(wplanes);
    __nv_bfloat16* w_lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(wplanes) + wp);
    launch_k(weight_planes_kernel, (int)std::min<long>((9L * Cout * Cin + 255) / 256, 1184), 256, 0, st, w, Cout, Cin, dgrad, fmt, w_hi, w_lo, scale2);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// activation planes given ([B][H][W][Kc] bf16 hi / lo); weight planes are rebuilt into wscratch (tiny)
int conv_tc_planes(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* out, float* stats,
                   int B, int H, int W, int Cin, int Cout, int dgrad, void* wscratch, cudaStream_t st) {
    const int rc = conv_tc_weight_planes(w, Cin, Cout, dgrad, wscratch, st);
    if (rc) return rc;
    return conv_tc_planes_w(a_hi, a_lo, wscratch, bias, out, stats, B, H, W, Cin, Cout, dgrad, st);
}

// same with the weight planes already built (conv_tc_weight_planes)
int conv_tc_planes_w(const void* a_hi, const void* a_lo, const void* wplanes, const float* bias, float* out, float* stats,
                     int B, int H, int W, int Cin, int Cout, int dgrad, cudaStream_t st, int terms, int fmt,
                     const float* out_scale, const float* out2_scale) {
    const int Kc = dgrad ? Cout : Cin, Nc = dgrad ? Cin : Cout;
    SED_REQUIRE(terms == 1 || terms == 2 || terms == 3, SEDB200_EINVAL, "conv_tc: terms = %d", terms);
    SED_REQUIRE(terms != 2 || (fmt == kPlaneF16 && out2_scale && a_lo), SEDB200_EINVAL, "conv_tc: the fp16 + fp8 mode needs fp16 planes, the correction planes and their scale");
    if (terms == 1 || !a_lo) a_lo = a_hi;                       // 1-term: the lo maps are encoded but never used
    SED_REQUIRE(conv_tc_supported(H, W, Kc, Nc), SEDB200_ESHAPE, "conv_tc: shape H=%d W=%d K=%d N=%d unsupported", H, W, Kc, Nc);
    const size_t wp = ((size_t)9 * Nc * Kc * 2 + 1023) & ~(size_t)1023;
    const __nv_bfloat16* w_hi = reinterpret_cast<const __nv_bfloat16*>(wplanes);
    const __nv_bfloat16* w_lo = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const char*>(wplanes) + wp);

    const int Ht = kTileM / W;
    ConvTcParams p;
    p.B = B; p.H = H; p.W = W; p.Ht = Ht;
    p.lgW = 0;
    while ((1 << p.lgW) < W) ++p.lgW;
    p.tiles_per_img = (H + Ht - 1) / Ht;
    p.n_tiles_n = Nc / kTileN;
    p.m_tiles = B * p.tiles_per_img;
    p.kchunks = Kc / kBlockK;
    p.n_total = Nc;
    p.terms = terms;
    // halo mode (one A box per tap COLUMN, conv_halo_mode) whenever a tap-row offset is a whole number of swizzle atoms.
    // Measured (profiles/README.md, "r02 halo boxes"): no gain while the MMA-issuing thread was the bound, conv 2 forward
    // 0.153 -> 0.134 ms and its data gradient 0.082 -> 0.073 ms once the issue stream was warp-uniform.
    // SEDB200_CONV_HALO=0 turns it off (the parity tests run both settings).
    const char* e_halo = std::getenv("SEDB200_CONV_HALO");
    bool halo = (W % 4 == 0) && !(e_halo && std::atoi(e_halo) == 0);
    // two M tiles per work item when the CTAs stay balanced with half as many items
    const char* e_tpi = std::getenv("SEDB200_CONV_TPI");
    const int force_tpi = e_tpi ? std::atoi(e_tpi) : 0;
    int tpi = 1;
    for (int pass = 0; pass < 2; ++pass) {
        const long pairs = halo ? (long)B * ((p.tiles_per_img + 1) / 2) * p.n_tiles_n : (long)((p.m_tiles + 1) / 2) * p.n_tiles_n;
        // (>= 3 items per CTA: at 3.5 pairs per CTA -- conv 3's data gradient at C2 -- the pair form still wins, 0.055 ->
        // 0.046 ms: half the weight bytes and half the per-tile hand-offs against one CTA round of imbalance)
        tpi = pairs >= 3L * sm_count() ? 2 : 1;
        if (force_tpi == 1 || force_tpi == 2) tpi = force_tpi;
        if (terms == 2) tpi = 1;                                 // the tile's second accumulator slot holds the fp8 pass
        if (halo && terms == 3) tpi = 1;                         // two A slots of two planes of two tiles leave no room for B
        if (!halo || conv_halo_mode(terms, tpi, W).n_b >= 2) break;
        halo = false;                                            // wide rows (W = 64, 128): the halo rows do not fit
    }
    int n_epi_warps, smem_bytes;
    p.halo = halo ? 1 : 0;
    p.tpi = tpi;
    p.planes = terms >= 2 ? 2 : 1;
    if (halo) {
        const HaloMode hm = conv_halo_mode(terms, tpi, W);
        p.groups_per_img = (p.tiles_per_img + tpi - 1) / tpi;
        p.n_a = hm.n_a; p.a_plane_bytes = hm.a_plane_bytes; p.a_slot_bytes = hm.a_slot_bytes; p.ring_bytes = hm.ring_bytes;
        p.stage_bytes = hm.b_slot_bytes; p.n_stages = hm.n_b;    // the B ring
        p.total_items = B * p.groups_per_img * p.n_tiles_n;
        n_epi_warps = hm.n_epi_warps; smem_bytes = hm.smem_bytes;
    } else {
        const ConvMode md = conv_mode(terms, tpi);
        SED_REQUIRE(md.n_stages >= 2, SEDB200_ESHAPE, "conv_tc: no room for two pipeline stages");
        p.groups_per_img = 0; p.n_a = 0; p.a_plane_bytes = 0; p.a_slot_bytes = 0; p.ring_bytes = 0;
        p.stage_bytes = md.stage_bytes; p.n_stages = md.n_stages;
        p.total_items = ((p.m_tiles + tpi - 1) / tpi) * p.n_tiles_n;
        n_epi_warps = md.n_epi_warps; smem_bytes = md.smem_bytes;
    }
    p.n_epi_warps = n_epi_warps;
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
    {
        const uint64_t dims[4] = {(uint64_t)Kc, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Kc * 2, (uint64_t)W * Kc * 2, (uint64_t)H * W * Kc * 2};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)W, (uint32_t)(halo ? tpi * Ht + 2 : Ht), 1};
        int rc = encode_tmap_bf16(&tmA_hi, a_hi, 4, dims, strides, box);
        if (rc) return rc;
        rc = encode_tmap_bf16(&tmA_lo, a_lo, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)Kc, (uint64_t)9 * Nc};
        const uint64_t strides[1] = {(uint64_t)Kc * 2};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)kTileN};
        int rc = encode_tmap_bf16(&tmB_hi, w_hi, 2, dims, strides, box);
        if (rc) return rc;
        rc = encode_tmap_bf16(&tmB_lo, w_lo, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.idesc = fmt == kPlaneF16 ? idesc_f16(kTileM, kTileN, 0, 0) : idesc_bf16(kTileM, kTileN, 0, 0);
    p.out_scale = out_scale;
    p.out2_scale = out2_scale;
    p.out = out; p.bias = bias; p.out_ld = Nc; p.stats = stats;
    { const int rc = ensure_dyn_smem((const void*)conv_tc_kernel, 227 * 1024); if (rc) return rc; }
    const int grid = std::min(p.total_items, sm_count());
    SED_REQUIRE(terms != 1 || !stats, SEDB200_EINVAL, "conv_tc: BatchNorm statistics need a forward mode");
    launch_k(conv_tc_kernel, grid, kConvThreads, smem_bytes, st, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int conv_tc_stat_tiles(int B, int H, int W) { return B * ((H + kTileM / W - 1) / (kTileM / W)); }

int conv_tc_forward(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin,
                    int Cout, int dgrad, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    // dgrad: `in` is dY [B,H,W,Cout]; the result is dX [B,H,W,Cin]
    const int Kc = dgrad ? Cout : Cin, Nc = dgrad ? Cin : Cout;
    SED_REQUIRE(conv_tc_supported(H, W, Kc, Nc), SEDB200_ESHAPE, "conv_tc: shape H=%d W=%d K=%d N=%d unsupported", H, W, Kc, Nc);
    SED_REQUIRE(scratch_bytes >= conv_tc_scratch_bytes(B, H, W, Kc, Nc), SEDB200_EWORKSPACE, "conv_tc: scratch too small");
    const size_t act = ((size_t)B * H * W * Kc * 2 + 1023) & ~(size_t)1023;
    char* s = reinterpret_cast<char*>(scratch);
    const long n4 = (long)B * H * W * Kc / 4;
    launch_k(split_planes_kernel, (int)std::min<long>((n4 + 255) / 256, 148L * 8), 256, 0, st,
        in, reinterpret_cast<__nv_bfloat16*>(s), reinterpret_cast<__nv_bfloat16*>(s + act), n4);
    SED_POST_LAUNCH();
    return conv_tc_planes(s, s + act, w, bias, out, nullptr, B, H, W, Cin, Cout, dgrad, s + 2 * act, st);
}


bool wgrad_tc_supported(int H, int W, int Cin, int Cout) {
    return Cin % 128 == 0 && Cout % 128 == 0 && W >= 1 && W <= 32 && (32 % W) == 0 && H >= 1;
}

static int wgrad_slices(int Cin, int Cout) {
    const int items = (Cout / 128) * (Cin / 128) * 3;
    return std::max(1, sm_count() / items);
}

size_t wgrad_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout) {
    const size_t px = (size_t)B * H * W;
    const size_t planes = 2 * ((px * Cout * 2 + 1023) & ~(size_t)1023) + 2 * ((px * Cin * 2 + 1023) & ~(size_t)1023);
    return planes + (size_t)wgrad_slices(Cin, Cout) * 9 * Cout * Cin * 4;
}

// dw[Cout][Cin][3][3] = wgrad(dy[B,H,W,Cout], in[B,H,W,Cin]); if the bf16 planes already exist they can be
// passed in (null -> they are produced here from the fp32 tensors)
size_t wgrad_tc_part_bytes(int Cin, int Cout) { return (size_t)wgrad_slices(Cin, Cout) * 9 * Cout * Cin * 4; }

// planes given: dY [B][H][W][Cout] and In [B][H][W][Cin] as bf16 hi / lo
int wgrad_tc_planes(const void* y_hi, const void* y_lo, const void* x_hi, const void* x_lo, float* dw, int B, int H,
                    int W, int Cin, int Cout, float* part, cudaStream_t st, int terms, int fmt, const float* out_scale,
                    int max_stages) {
    SED_REQUIRE(wgrad_tc_supported(H, W, Cin, Cout), SEDB200_ESHAPE, "wgrad_tc: shape H=%d W=%d Cin=%d Cout=%d unsupported", H, W, Cin, Cout);
    SED_REQUIRE(terms == 1 || terms == 3, SEDB200_EINVAL, "wgrad_tc: terms = %d", terms);
    if (terms == 1 || !y_lo) y_lo = y_hi;
    if (terms == 1 || !x_lo) x_lo = x_hi;
    const int Hk = kWgKp / W;
    CUtensorMap tmY_hi, tmY_lo, tmX_hi, tmX_lo;
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)Hk, 1};
    {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
        int rc = encode_tmap_bf16(&tmY_hi, y_hi, 4, dims, strides, box);
        if (rc) return rc;
        rc = encode_tmap_bf16(&tmY_lo, y_lo, 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        int rc = encode_tmap_bf16(&tmX_hi, x_hi, 4, dims, strides, box);
        if (rc) return rc;
        rc = encode_tmap_bf16(&tmX_lo, x_lo, 4, dims, strides, box);
        if (rc) return rc;
    }
    WgradTcParams p;
    p.B = B; p.H = H; p.W = W; p.Hk = Hk;
    p.hblocks_per_img = (H + Hk - 1) / Hk;
    p.total_kblocks = B * p.hblocks_per_img;
    p.n_mt = Cout / 128; p.n_nt = Cin / 128;
    p.slices = std::min(wgrad_slices(Cin, Cout), p.total_kblocks);
    p.part = part; p.Cout = Cout; p.Cin = Cin; p.terms = terms;
    p.idesc = fmt == kPlaneF16 ? idesc_f16(128, 128, 1, 1) : idesc_bf16(128, 128, 1, 1);
    const char* e_halo = std::getenv("SEDB200_WGRAD_HALO");
    p.halo = (terms == 1 && W % 4 == 0 && !(e_halo && std::atoi(e_halo) == 0)) ? 1 : 0;
    p.xh_bytes = (kWgKp + 2 * W) * 128;                             // one channel half of the In box: Hk + 2 image rows
    p.stage_bytes = p.halo ? 2 * kWgBox + 2 * p.xh_bytes : (terms == 1 ? kWgStageBytes / 2 : kWgStageBytes);
    if (p.halo) {                                                   // the In box of Hk + 2 rows (tmX_lo is free in this mode)
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t hbox[4] = {64, (uint32_t)W, (uint32_t)(Hk + 2), 1};
        const int rc = encode_tmap_bf16(&tmX_lo, x_hi, 4, dims, strides, hbox);
        if (rc) return rc;
    }
    // pipeline depth: 6 half-size stages (single pass) / 3 (3-term) fill the SM; a caller that wants the kernel to
    // share its SMs with another one (crnn_backward_impl: beside the BatchNorm / block-0 backward) passes max_stages =
    // the shared memory it may take in units of the 32 KB single-pass stage (the smaller halo stages: more of them)
    const int full_stages = terms == 1 ? kMaxStages : kStages;
    const int legacy_stage = terms == 1 ? kWgStageBytes / 2 : kWgStageBytes;
    int budget = full_stages * legacy_stage;
    if (max_stages >= 2 && max_stages < full_stages) budget = max_stages * legacy_stage;
    p.n_stages = std::max(2, std::min(kMaxStages, budget / p.stage_bytes));
    const int smem_bytes = p.n_stages * p.stage_bytes + 1024 + 256;
    { const int rc = ensure_dyn_smem((const void*)wgrad_tc_kernel, kWgSmemBytes); if (rc) return rc; }
    const int grid = p.n_mt * p.n_nt * 3 * p.slices;
    launch_k(wgrad_tc_kernel, grid, kThreads, smem_bytes, st, tmY_hi, tmY_lo, tmX_hi, tmX_lo, p);
    SED_POST_LAUNCH();
    launch_k(wgrad_reduce_kernel, (int)std::min<long>((9L * Cout * Cin + 255) / 256, 1184), 256, 0, st, part, p.slices, Cout, Cin, dw, out_scale);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// dw[Cout][Cin][3][3] = wgrad(dy[B,H,W,Cout], in[B,H,W,Cin]) from fp32 tensors (planes are produced here)
int wgrad_tc(const float* dy, const float* in, float* dw, int B, int H, int W, int Cin, int Cout, void* scratch,
             size_t scratch_bytes, cudaStream_t st) {
    SED_REQUIRE(wgrad_tc_supported(H, W, Cin, Cout), SEDB200_ESHAPE, "wgrad_tc: shape H=%d W=%d Cin=%d Cout=%d unsupported", H, W, Cin, Cout);
    SED_REQUIRE(scratch_bytes >= wgrad_tc_scratch_bytes(B, H, W, Cin, Cout), SEDB200_EWORKSPACE, "wgrad_tc: scratch too small");
    const size_t px = (size_t)B * H * W;
    const size_t ysz = (px * Cout * 2 + 1023) & ~(size_t)1023, xsz = (px * Cin * 2 + 1023) & ~(size_t)1023;
    char* s = reinterpret_cast<char*>(scratch);
    long n4 = (long)px * Cout / 4;
    launch_k(split_planes_kernel, (int)std::min<long>((n4 + 255) / 256, 148L * 8), 256, 0, st,
        dy, reinterpret_cast<__nv_bfloat16*>(s), reinterpret_cast<__nv_bfloat16*>(s + ysz), n4);
    SED_POST_LAUNCH();
    n4 = (long)px * Cin / 4;
    launch_k(split_planes_kernel, (int)std::min<long>((n4 + 255) / 256, 148L * 8), 256, 0, st,
        in, reinterpret_cast<__nv_bfloat16*>(s + 2 * ysz), reinterpret_cast<__nv_bfloat16*>(s + 2 * ysz + xsz), n4);
    SED_POST_LAUNCH();
    return wgrad_tc_planes(s, s + ysz, s + 2 * ysz, s + 2 * ysz + xsz, dw, B, H, W, Cin, Cout,
                           reinterpret_cast<float*>(s + 2 * ysz + 2 * xsz), st);
}

}  // namespace sedb200

using namespace sedb200;

extern "C" {

size_t sedb200_conv3x3_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout) {
    return std::max(conv_tc_scratch_bytes(B, H, W, Cin, Cout), conv_tc_scratch_bytes(B, H, W, Cout, Cin));
}

int sedb200_conv3x3_tc(const float* in_dev, const float* weight_dev, const float* bias_dev, float* out_dev, int B,
                       int H, int W, int Cin, int Cout, int dgrad, void* scratch_dev, size_t scratch_bytes,
                       void* stream) {
    SED_REQUIRE(in_dev && weight_dev && out_dev && scratch_dev, SEDB200_EINVAL, "conv3x3_tc: null buffer");
    SED_REQUIRE(B >= 1, SEDB200_EINVAL, "conv3x3_tc: batch %d", B);
    int rc = require_sm100();
    if (rc) return rc;
    return conv_tc_forward(in_dev, weight_dev, bias_dev, out_dev, B, H, W, Cin, Cout, dgrad, scratch_dev,
                           scratch_bytes, as_stream(stream));
}

size_t sedb200_conv3x3_wgrad_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout) {
    return wgrad_tc_scratch_bytes(B, H, W, Cin, Cout);
}

int sedb200_conv3x3_wgrad_tc(const float* dy_dev, const float* in_dev, float* dw_dev, int B, int H, int W, int Cin,
                             int Cout, void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(dy_dev && in_dev && dw_dev && scratch_dev, SEDB200_EINVAL, "conv3x3_wgrad_tc: null buffer");
    SED_REQUIRE(B >= 1, SEDB200_EINVAL, "conv3x3_wgrad_tc: batch %d", B);
    int rc = require_sm100();
    if (rc) return rc;
    return wgrad_tc(dy_dev, in_dev, dw_dev, B, H, W, Cin, Cout, scratch_dev, scratch_bytes, as_stream(stream));
}

}  // extern "C"

// ---------------------------------------------------------------------------------- plane-native test hook
// The CRNN flow never calls the fp32 entry points above: its conv blocks run conv_tc_planes_w / wgrad_tc_planes on fp16
// planes (forward: fp16 pass + e4m3 correction pass; gradients: one fp16 pass; halo boxes; tile pairs).  This hook runs
// exactly those launches on fp32 tensors -- the planes are produced here with the SAME helpers the pool kernels use
// (crnn_block.cuh) -- so that tests can hold them against a float64 convolution tap by tap.
namespace sedb200 {
namespace {
__global__ void __launch_bounds__(256)
test_planes_kernel(const float* __restrict__ x, long n_pix, int C4, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ c8, int with_c8) {
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * C4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const long pix = i / C4;
        const int c4 = (int)(i - pix * C4);
        if (with_c8) store_planes4(hi, c8, pix, c4, C4, v);
        else store_dy4(hi, i, v, 1.0f);
    }
}
__global__ void test_one_kernel(float* p) { pdl_wait(); p[0] = 1.0f; p[1] = 1.0f; }
}  // namespace
}  // namespace sedb200

extern "C" {

size_t sedb200_conv3x3_planes_test_scratch_bytes(int B, int H, int W, int Cin, int Cout) {
    const size_t px = (size_t)B * H * W, cmax = (size_t)std::max(Cin, Cout);
    const size_t plane = (px * cmax * 2 + 1023) & ~(size_t)1023;
    return 4 * plane + conv_tc_weight_scratch_bytes(Cin, Cout) + wgrad_tc_part_bytes(Cin, Cout) + 4096;
}

// mode 0: out [B][H][W][Cout] = conv(in [B][H][W][Cin], w)                        forward, fp16 + e4m3 passes
// mode 1: out [B][H][W][Cin]  = conv_transposed(in = dY [B][H][W][Cout], w)       data gradient, one fp16 pass
// mode 2: out [Cout][Cin][3][3] = wgrad(in = dY [B][H][W][Cout], in2 = x [B][H][W][Cin])   weight gradient, one pass
int sedb200_conv3x3_planes_test(const float* in_dev, const float* in2_dev, const float* weight_dev, float* out_dev, int B,
                                int H, int W, int Cin, int Cout, int mode, void* scratch_dev, size_t scratch_bytes,
                                void* stream) {
    SED_REQUIRE(in_dev && out_dev && scratch_dev && (mode == 2 ? in2_dev != nullptr : weight_dev != nullptr), SEDB200_EINVAL,
                "conv3x3_planes_test: null buffer");
    SED_REQUIRE(mode >= 0 && mode <= 2 && B >= 1, SEDB200_EINVAL, "conv3x3_planes_test: mode %d batch %d", mode, B);
    SED_REQUIRE(scratch_bytes >= sedb200_conv3x3_planes_test_scratch_bytes(B, H, W, Cin, Cout), SEDB200_EWORKSPACE,
                "conv3x3_planes_test: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const long px = (long)B * H * W;
    const size_t cmax = (size_t)std::max(Cin, Cout), plane = ((size_t)px * cmax * 2 + 1023) & ~(size_t)1023;
    char* s = reinterpret_cast<char*>(scratch_dev);
    __nv_bfloat16 *p0 = reinterpret_cast<__nv_bfloat16*>(s), *p1 = reinterpret_cast<__nv_bfloat16*>(s + plane);
    __nv_bfloat16 *p2 = reinterpret_cast<__nv_bfloat16*>(s + 2 * plane), *p3 = reinterpret_cast<__nv_bfloat16*>(s + 3 * plane);
    char* wpl = s + 4 * plane;
    float* wpart = reinterpret_cast<float*>(wpl + conv_tc_weight_scratch_bytes(Cin, Cout));
    float* one = reinterpret_cast<float*>(reinterpret_cast<char*>(wpart) + wgrad_tc_part_bytes(Cin, Cout));
    launch_k(test_one_kernel, 1, 1, 0, st, one);
    SED_POST_LAUNCH();
    auto planes = [&](const float* x, int C, __nv_bfloat16* hi, __nv_bfloat16* c8, int with_c8) {
        const long n = px * (C / 4);
        launch_k(test_planes_kernel, (int)std::min<long>((n + 255) / 256, 148L * 8), 256, 0, st, x, px, C / 4, hi, c8, with_c8);
    };
    if (mode == 0) {
        SED_REQUIRE(conv_tc_supported(H, W, Cin, Cout), SEDB200_ESHAPE, "conv3x3_planes_test: forward shape unsupported");
        planes(in_dev, Cin, p0, p1, 1);
        SED_POST_LAUNCH();
        float* sc2 = one + 2;
        rc = conv_tc_weight_planes(weight_dev, Cin, Cout, 0, wpl, st, kPlaneF16, sc2);
        if (rc) return rc;
        return conv_tc_planes_w(p0, p1, wpl, nullptr, out_dev, nullptr, B, H, W, Cin, Cout, 0, st, 2, kPlaneF16, nullptr, sc2 + 1);
    }
    if (mode == 1) {
        SED_REQUIRE(conv_tc_supported(H, W, Cout, Cin), SEDB200_ESHAPE, "conv3x3_planes_test: data-gradient shape unsupported");
        planes(in_dev, Cout, p0, nullptr, 0);
        SED_POST_LAUNCH();
        rc = conv_tc_weight_planes(weight_dev, Cin, Cout, 1, wpl, st, kPlaneF16);
        if (rc) return rc;
        return conv_tc_planes_w(p0, nullptr, wpl, nullptr, out_dev, nullptr, B, H, W, Cin, Cout, 1, st, 1, kPlaneF16, one);
    }
    SED_REQUIRE(wgrad_tc_supported(H, W, Cin, Cout), SEDB200_ESHAPE, "conv3x3_planes_test: weight-gradient shape unsupported");
    planes(in_dev, Cout, p0, nullptr, 0);
    SED_POST_LAUNCH();
    planes(in2_dev, Cin, p2, p3, 1);
    SED_POST_LAUNCH();
    return wgrad_tc_planes(p0, nullptr, p2, nullptr, out_dev, B, H, W, Cin, Cout, wpart, st, 1, kPlaneF16, one);
}

}  // extern "C"
