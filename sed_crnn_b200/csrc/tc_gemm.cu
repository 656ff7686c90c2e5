// tc_gemm.cu -- plain GEMMs on tcgen05 with the same fp32-grade 3-term bf16 split as tc_conv.cu.
//
//   D[M][N] (+bias) = sum_k A(m,k) * B(n,k)
//
// Each operand is either K-major (memory [rows = M or N][K], k contiguous) or MN-major (memory [K][rows],
// the M/N index contiguous), so the three shapes of the GRU / dense part of the CRNN map on one kernel:
//   gi   = X W_ih^T + b_ih        A K-major  [B*T][in],   B K-major  [6H][in]        (nn.GRU input projection)
//   dX   = dgi W_ih               A K-major  [B*T][6H],   B MN-major [6H][in]
//   dW   = dgi^T X                A MN-major [B*T][6H],   B MN-major [B*T][in]       (reduction over B*T rows,
//                                                                                     split-K + fixed-order reduce)
// Tile 128 x 128 x 64, 3-stage TMA ring, two TMEM accumulators, same warp roles as conv_tc_kernel.
#include "tc_umma.cuh"
#include "tc_gemm.cuh"
#include "gemm_simt.cuh"

#include <algorithm>

namespace sedb200 {
namespace {
using namespace umma;

constexpr int kStages = 3;
constexpr int kTile = 128, kBlockK = 64;
constexpr int kTileBytes = kTile * kBlockK * 2;          // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;
constexpr int kEpiPitch = 36;                            // floats per staged row: 16 B aligned, conflict-free as LDS/STS.128
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;        // one 32 x 32 staging tile per epilogue warp
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpiBytes;
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 256;

struct GemmTcParams {
    int M, N, n_mt, n_nt, kb_total, slices, n_items;
    float* out;              // slices == 1: [M][out_ld] (+bias); else partials [slice][M][N]
    const float* bias;
    long out_ld;
};

template <bool MN>
__device__ __forceinline__ void load_operand(unsigned char* dst, const CUtensorMap* tm, uint64_t* bar, int tile, int kb) {
    if (MN) {                // tensor dims {rows(M|N), K}: two {64 x 64} boxes
        tma_load_2d(dst, tm, bar, tile * kTile, kb * kBlockK);
        tma_load_2d(dst + kTileBytes / 2, tm, bar, tile * kTile + 64, kb * kBlockK);
    } else {                 // tensor dims {K, rows}: one {64 k x 128 rows} box
        tma_load_2d(dst, tm, bar, kb * kBlockK, tile * kTile);
    }
}
template <bool MN>
__device__ __forceinline__ uint64_t operand_desc(uint32_t addr, int k) {
    return MN ? smem_desc_sw128(addr + k * 2048, kTileBytes / 2, 1024) : smem_desc_sw128(addr + k * 32, 16, 1024);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const GemmTcParams p) {
    pdl_wait();
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float* epi_stage = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int kb_per = (p.kb_total + p.slices - 1) / p.slices;

    // item -> (slice, nt, mt)
    auto decode = [&](int item, int& mt, int& nt, int& kb0, int& kb1, int& slice) {
        slice = item % p.slices;
        const int t = item / p.slices;
        nt = t % p.n_nt;
        mt = t / p.n_nt;
        kb0 = slice * kb_per;
        kb1 = min(p.kb_total, kb0 + kb_per);
    };

    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int mt, nt, kb0, kb1, slice;
            decode(item, mt, nt, kb0, kb1, slice);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty + stage, phase ^ 1);
                unsigned char* st = smem + stage * kStageBytes;
                mbar_expect_tx(full + stage, kStageBytes);
                load_operand<A_MN>(st, &tmA_hi, full + stage, mt, kb);
                load_operand<A_MN>(st + kTileBytes, &tmA_lo, full + stage, mt, kb);
                load_operand<B_MN>(st + 2 * kTileBytes, &tmB_hi, full + stage, nt, kb);
                load_operand<B_MN>(st + 3 * kTileBytes, &tmB_lo, full + stage, nt, kb);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // warp-wide control flow, one elected lane issues (tc_umma.cuh: elect_one): descriptors in uniform registers
        constexpr uint32_t idesc = idesc_bf16(kTile, kTile, A_MN ? 1 : 0, B_MN ? 1 : 0);
        const uint64_t da0 = operand_desc<A_MN>(smem_u32(smem), 0), db0 = operand_desc<B_MN>(smem_u32(smem), 0);
        constexpr uint32_t stage_u = kStageBytes >> 4, tile_u = kTileBytes >> 4;
        constexpr uint32_t ka_u = (A_MN ? 2048 : 32) >> 4, kb_u = (B_MN ? 2048 : 32) >> 4;
        int stage = 0; uint32_t phase = 0;
        int buf = 0; uint32_t bphase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int mt, nt, kb0, kb1, slice;
            decode(item, mt, nt, kb0, kb1, slice);
            if (kb1 <= kb0) continue;                      // empty slice: the epilogue writes zeros itself
            mbar_wait(tempty + buf, bphase ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + buf * kTile;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dah0 = da0 + (uint64_t)(stage * stage_u), dal0 = dah0 + tile_u;
                    const uint64_t dbh0 = db0 + (uint64_t)(stage * stage_u + 2 * tile_u), dbl0 = dbh0 + tile_u;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        mma_bf16(d, dah0 + k * ka_u, dbh0 + k * kb_u, idesc, (kb != kb0 || k != 0));
                        mma_bf16(d, dah0 + k * ka_u, dbl0 + k * kb_u, idesc, 1);
                        mma_bf16(d, dal0 + k * ka_u, dbh0 + k * kb_u, idesc, 1);
                    }
                    mma_commit(empty + stage);
                    if (kb == kb1 - 1) mma_commit(tfull + buf);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        int buf = 0; uint32_t bphase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int mt, nt, kb0, kb1, slice;
            decode(item, mt, nt, kb0, kb1, slice);
            const bool has = kb1 > kb0;
            if (has) {
                mbar_wait(tfull + buf, bphase);
                tc_fence_after();
            }
            // TMEM hands every lane one ROW of the tile; storing that directly would scatter 16 B pieces over 32 rows
            // per instruction.  Each 32 x 32 chunk is transposed through a per-warp staging tile instead, so that one
            // store instruction writes four full 128 B row segments.
            const int row0 = mt * kTile + q * 32;
            float* dst0 = p.slices == 1 ? p.out + (long)row0 * p.out_ld : p.out + ((long)slice * p.M + row0) * p.N;
            const long ld = p.slices == 1 ? p.out_ld : (long)p.N;
            float* stg = epi_stage + q * (32 * kEpiPitch);
            const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
#pragma unroll 1
            for (int cc = 0; cc < kTile / 32; ++cc) {
                const int col0 = nt * kTile + cc * 32;
                if (col0 >= p.N) break;                     // tile columns beyond N (N % 4 == 0)
                float v[32];
                if (has) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * kTile + cc * 32, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * kEpiPitch + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                const int col = col0 + sub_c;
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && col < p.N) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    float4 o = *reinterpret_cast<const float4*>(stg + (r + sub_r) * kEpiPitch + sub_c);
                    o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    if (row0 + r + sub_r < p.M && col < p.N)
                        *reinterpret_cast<float4*>(dst0 + (long)(r + sub_r) * ld + col) = o;
                }
                __syncwarp();
            }
            if (has) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + buf);
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long n4) {
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        __nv_bfloat16 h[4], l[4];
        split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
}
// h_{t-1} of both GRU directions as one [B*T][2H] matrix: columns [0,H) = out[b][t-1][0:H] (0 at t = 0),
// columns [H,2H) = out[b][t+1][H:2H] (0 at t = T-1); written directly as bf16 planes
__global__ void __launch_bounds__(256)
hprev_planes_kernel(const float* __restrict__ out, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                    long n, int T, int H) {
    pdl_wait();
    const int H2 = 2 * H;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % H2);
        const long row = i / H2;
        const int t = (int)(row % T);
        const int tp = c < H ? t - 1 : t + 1;
        const float v = (tp >= 0 && tp < T) ? __ldg(out + (row - t + tp) * H2 + c) : 0.0f;
        split1(v, hi[i], lo[i]);
    }
}

inline size_t plane_bytes(long rows, long cols) { return ((size_t)rows * cols * 2 + 1023) & ~(size_t)1023; }

template <bool A_MN, bool B_MN>
int launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
           const GemmTcParams& p, cudaStream_t st) {
    { const int rc = ensure_dyn_smem((const void*)gemm_tc_kernel<A_MN, B_MN>, kSmemBytes); if (rc) return rc; }
    const int grid = std::min(p.n_items, sm_count());
    launch_k(gemm_tc_kernel<A_MN, B_MN>, grid, kThreads, kSmemBytes, st, a_hi, a_lo, b_hi, b_lo, p);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// tensor map of one operand plane: K-major -> dims {K, rows}, box {64, 128}; MN-major -> dims {rows, K}, box {64, 64}
int operand_tmap(CUtensorMap* tm, const void* plane, bool mn_major, long rows, long K) {
    if (mn_major) {
        const uint64_t dims[2] = {(uint64_t)rows, (uint64_t)K};
        const uint64_t strides[1] = {(uint64_t)rows * 2};
        const uint32_t box[2] = {64, 64};
        return encode_tmap_bf16(tm, plane, 2, dims, strides, box);
    }
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, 128};
    return encode_tmap_bf16(tm, plane, 2, dims, strides, box);
}

}  // namespace

bool gemm_tc_supported(int M, int N, int K) { return M >= 1 && N >= 4 && N % 8 == 0 && M % 8 == 0 && K % 8 == 0 && K >= 8; }

size_t gemm_tc_scratch_bytes(int M, int N, int K, int a_mn, int b_mn, int split_k) {
    (void)a_mn; (void)b_mn;
    size_t s = 2 * plane_bytes(M, K) + 2 * plane_bytes(N, K);
    if (split_k) s += (size_t)sm_count() * M * N * 4 + 1024;
    return s;
}

int split_planes(const float* x, void* hi, void* lo, long n, cudaStream_t st) {
    SED_REQUIRE(n % 4 == 0, SEDB200_ESHAPE, "split_planes: n %% 4 != 0");
    const long n4 = n / 4;
    launch_k(split_kernel, (int)std::min<long>((n4 + 255) / 256, 148L * 8), 256, 0, st,
        x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n4);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int hprev_planes(const float* out, void* hi, void* lo, long rows, int T, int H, cudaStream_t st) {
    const long n = rows * 2 * H;
    launch_k(hprev_planes_kernel, (int)std::min<long>((n + 255) / 256, 148L * 8), 256, 0, st,
        out, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n, T, H);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// Operands are given as ready bf16 planes (a_hi/a_lo, b_hi/b_lo).  a_mn / b_mn: 0 = K-major ([rows][K]),
// 1 = MN-major ([K][rows]).  split_k: 0 -> out[M][out_ld] (+bias); 1 -> K is split over the SMs, partials go
// to `part` ([slices][M][N]) and are reduced into out[M][N] (contiguous) in a fixed order.
int gemm_tc(const void* a_hi, const void* a_lo, int a_mn, const void* b_hi, const void* b_lo, int b_mn, int M, int N,
            int K, const float* bias, float* out, long out_ld, int split_k, float* part, cudaStream_t st) {
    SED_REQUIRE(gemm_tc_supported(M, N, K), SEDB200_ESHAPE, "gemm_tc: M=%d N=%d K=%d unsupported", M, N, K);
    CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
    int rc = operand_tmap(&tA_hi, a_hi, a_mn, M, K);
    if (rc) return rc;
    rc = operand_tmap(&tA_lo, a_lo, a_mn, M, K);
    if (rc) return rc;
    rc = operand_tmap(&tB_hi, b_hi, b_mn, N, K);
    if (rc) return rc;
    rc = operand_tmap(&tB_lo, b_lo, b_mn, N, K);
    if (rc) return rc;
    GemmTcParams p;
    p.M = M; p.N = N;
    p.n_mt = (M + kTile - 1) / kTile;
    p.n_nt = (N + kTile - 1) / kTile;
    p.kb_total = (K + kBlockK - 1) / kBlockK;
    const int tiles = p.n_mt * p.n_nt;
    p.slices = split_k ? std::max(1, std::min(p.kb_total, sm_count() / tiles)) : 1;
    p.n_items = tiles * p.slices;
    p.bias = split_k ? nullptr : bias;
    p.out = p.slices == 1 ? out : part;
    p.out_ld = p.slices == 1 ? out_ld : N;
    if (a_mn && b_mn) rc = launch<true, true>(tA_hi, tA_lo, tB_hi, tB_lo, p, st);
    else if (!a_mn && b_mn) rc = launch<false, true>(tA_hi, tA_lo, tB_hi, tB_lo, p, st);
    else if (!a_mn && !b_mn) rc = launch<false, false>(tA_hi, tA_lo, tB_hi, tB_lo, p, st);
    else return fail(SEDB200_ESHAPE, "gemm_tc: A MN-major with B K-major is not instantiated");
    if (rc) return rc;
    if (p.slices > 1) return reduce_partials(part, out, (long)M * N, p.slices, st);
    return SEDB200_OK;
}

}  // namespace sedb200

using namespace sedb200;

extern "C" {

size_t sedb200_gemm_tc_scratch_bytes(int M, int N, int K) { return gemm_tc_scratch_bytes(M, N, K, 0, 0, 1); }

// test entry: fp32 operands in their natural row-major layouts, split into planes here
int sedb200_gemm_tc(const float* a_dev, int a_mn, const float* b_dev, int b_mn, int M, int N, int K,
                    const float* bias_dev, float* out_dev, int split_k, void* scratch_dev, size_t scratch_bytes,
                    void* stream) {
    SED_REQUIRE(a_dev && b_dev && out_dev && scratch_dev, SEDB200_EINVAL, "gemm_tc: null buffer");
    SED_REQUIRE(scratch_bytes >= sedb200_gemm_tc_scratch_bytes(M, N, K), SEDB200_EWORKSPACE, "gemm_tc: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    char* s = reinterpret_cast<char*>(scratch_dev);
    const size_t pa = ((size_t)M * K * 2 + 1023) & ~(size_t)1023, pb = ((size_t)N * K * 2 + 1023) & ~(size_t)1023;
    void *a_hi = s, *a_lo = s + pa, *b_hi = s + 2 * pa, *b_lo = s + 2 * pa + pb;
    float* part = reinterpret_cast<float*>(s + 2 * pa + 2 * pb);
    rc = split_planes(a_dev, a_hi, a_lo, (long)M * K, st);
    if (rc) return rc;
    rc = split_planes(b_dev, b_hi, b_lo, (long)N * K, st);
    if (rc) return rc;
    return gemm_tc(a_hi, a_lo, a_mn, b_hi, b_lo, b_mn, M, N, K, bias_dev, out_dev, N, split_k, part, st);
}

}  // extern "C"
