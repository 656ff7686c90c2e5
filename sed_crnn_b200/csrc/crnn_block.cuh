// crnn_block.cuh -- device helpers shared by the conv-block kernels (crnn.cu, conv0_lean.cu): the counter-based
// dropout generator, the bf16 hi / lo plane store, the geometry of a BN + ReLU + max-pool block.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "common.cuh"

namespace sedb200 {

constexpr int kC0Rows = 8;             // image rows per row group of the direct first-block kernels
inline bool conv0_direct_ok(int cin, int C) { return (cin == 1 || cin == 2) && C % 128 == 0; }

// ----------------------------------------------------------------------------- dropout generator
// Two 32-bit hashes (murmur3 finaliser over the element-group index and the seed) per group of FOUR consecutive
// elements; element q of the group keeps its value iff the q-th 16-bit field is >= p * 65536.  Forward and backward
// kernels (and sedb200_crnn_dropout_mask) all go through dropout_keep4(seed, group index), so they agree on the mask
// without storing it.  32-bit arithmetic on purpose: the 64-bit splitmix of round 1 cost ~30 instructions per group,
// a tenth of the fused block-0 epilogue.
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
struct Keep4 { bool k[4]; };
__device__ __forceinline__ Keep4 dropout_keep4(unsigned long long seed, unsigned long long group, float p) {
    const unsigned lo = (unsigned)group, hi = (unsigned)(group >> 32);
    const unsigned a = mix32(lo * 0x9E3779B1u + hi * 0x7FEB352Du + (unsigned)seed);
    const unsigned b = mix32((lo ^ 0x68E31DA4u) * 0xB5297A4Du + hi + (unsigned)(seed >> 32));
    const unsigned thr = (unsigned)(p * 65536.0f);
    Keep4 r;
    r.k[0] = (a & 0xFFFFu) >= thr; r.k[1] = (a >> 16) >= thr;
    r.k[2] = (b & 0xFFFFu) >= thr; r.k[3] = (b >> 16) >= thr;
    return r;
}
__host__ __device__ inline unsigned long long block_seed(unsigned long long seed, int block) {
    return seed * 0x2545F4914F6CDD1Dull + (unsigned long long)(block + 1) * 0xD6E8FEB86659FD93ull;
}

// Activation planes of the plane-native conv blocks (what the tensor-core forward reads):
//   hi  [pixel][C] fp16           hi = fp16(x), saturating
//   c8  [pixel][C/64][2][64] e4m3 the two CORRECTION operands, 128 bytes per 64 channels: first e4m3(x), then
//                                 e4m3(2^12 (x - hi)) -- the same bytes per pixel as an fp16 plane, so that a 64-channel
//                                 K-block of it is one 128-byte-swizzled tile whose K = 128 fp8 contraction against the
//                                 weight tile laid out in the OPPOSITE order yields  x8 . w_lo8 + x_lo8 . w8  in one go.
// x ~ hi + 2^-12 * decode(lo8): 15 significand bits for the reader that needs the value back (bn_bwd_sums_act_kernel).
// The scales are static: the planes hold BatchNorm -> ReLU -> max-pool -> dropout outputs, O(1) by construction; e4m3
// covers 2^-9 .. 448 after scaling, values below flush (their correction term is below fp32 noise), values above
// saturate.  Storage type of all planes is `__nv_bfloat16` / bytes; what they mean is this contract.
constexpr float kActLo8Scale = 4096.0f;                    // (the e4m3 copy of x itself is unscaled: one multiply less per element)
// {fp16(a), fp16(b)} packed (a in the low half), round-to-nearest, saturating at +-65504: one F2FP instruction
__device__ __forceinline__ unsigned pack_half2_sat(float a, float b) {
    unsigned d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ float2 unpack_half2(unsigned v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
// {e4m3(a), e4m3(b)} packed (a in the low byte), saturating
__device__ __forceinline__ unsigned pack_e4m3x2(float a, float b) {
    unsigned short d;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ long c8_offset(long pix, int c4, int C4) {          // byte offset of channels 4*c4 .. +3
    return pix * 8 * C4 + (c4 >> 4) * 128 + (c4 & 15) * 4;
}
__device__ __forceinline__ void store_planes4(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ c8, long pix,
                                              int c4, int C4, const float4 v) {
    const unsigned h01 = pack_half2_sat(v.x, v.y), h23 = pack_half2_sat(v.z, v.w);
    const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
    reinterpret_cast<uint2*>(hi)[pix * C4 + c4] = make_uint2(h01, h23);
    unsigned char* p8 = reinterpret_cast<unsigned char*>(c8) + c8_offset(pix, c4, C4);
    *reinterpret_cast<unsigned*>(p8) = pack_e4m3x2(v.x, v.y) | (pack_e4m3x2(v.z, v.w) << 16);
    *reinterpret_cast<unsigned*>(p8 + 64) =
        pack_e4m3x2((v.x - f01.x) * kActLo8Scale, (v.y - f01.y) * kActLo8Scale) |
        (pack_e4m3x2((v.z - f23.x) * kActLo8Scale, (v.w - f23.y) * kActLo8Scale) << 16);
}
// eight consecutive channels (c4 even): one 16-byte store to the fp16 plane, two 8-byte stores to the e4m3 plane
__device__ __forceinline__ void store_planes8(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ c8, long pix,
                                              int c4, int C4, const float4 a, const float4 b) {
    const unsigned h0 = pack_half2_sat(a.x, a.y), h1 = pack_half2_sat(a.z, a.w), h2 = pack_half2_sat(b.x, b.y), h3 = pack_half2_sat(b.z, b.w);
    const float2 f0 = unpack_half2(h0), f1 = unpack_half2(h1), f2 = unpack_half2(h2), f3 = unpack_half2(h3);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(hi) + pix * C4 + c4) = make_uint4(h0, h1, h2, h3);
    unsigned char* p8 = reinterpret_cast<unsigned char*>(c8) + c8_offset(pix, c4, C4);
    *reinterpret_cast<uint2*>(p8) = make_uint2(pack_e4m3x2(a.x, a.y) | (pack_e4m3x2(a.z, a.w) << 16),
                                               pack_e4m3x2(b.x, b.y) | (pack_e4m3x2(b.z, b.w) << 16));
    constexpr float s = kActLo8Scale;
    *reinterpret_cast<uint2*>(p8 + 64) =
        make_uint2(pack_e4m3x2((a.x - f0.x) * s, (a.y - f0.y) * s) | (pack_e4m3x2((a.z - f1.x) * s, (a.w - f1.y) * s) << 16),
                   pack_e4m3x2((b.x - f2.x) * s, (b.y - f2.y) * s) | (pack_e4m3x2((b.z - f3.x) * s, (b.w - f3.y) * s) << 16));
}
__device__ __forceinline__ float4 load_planes4(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ c8,
                                               long pix, int c4, int C4) {
    const uint2 hb = __ldg(reinterpret_cast<const uint2*>(hi) + pix * C4 + c4);
    const unsigned lb = __ldg(reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(c8) + c8_offset(pix, c4, C4) + 64));
    const float2 f01 = unpack_half2(hb.x), f23 = unpack_half2(hb.y);
    __half2_raw l01 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(lb & 0xFFFFu), __NV_E4M3);
    __half2_raw l23 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(lb >> 16), __NV_E4M3);
    const float2 g01 = __half22float2(*reinterpret_cast<__half2*>(&l01)), g23 = __half22float2(*reinterpret_cast<__half2*>(&l23));
    constexpr float inv = 1.0f / kActLo8Scale;
    return make_float4(fmaf(g01.x, inv, f01.x), fmaf(g01.y, inv, f01.y), fmaf(g23.x, inv, f23.x), fmaf(g23.y, inv, f23.y));
}
// Gradient plane of a conv output: ONE fp16 plane of dy * scale, scale a power of two chosen per tensor from a bound
// on |dy| (dy_scale_kernel) so that the largest elements sit near 2^13; the contractions that consume it multiply
// their result by 1 / scale.  (The data / weight gradient contractions run single-pass: measured on the oracle,
// fp16-rounded gradient operands change the probabilities after a step by < 1e-4, the forward needs the split.)
__device__ __forceinline__ void store_dy4(__nv_bfloat16* __restrict__ hi, long i4, const float4 v, float scale) {
    reinterpret_cast<uint2*>(hi)[i4] = make_uint2(pack_half2_sat(v.x * scale, v.y * scale), pack_half2_sat(v.z * scale, v.w * scale));
}

struct PoolGeom {
    int H, W, Wo, C, p;
    long oB, oH, oW, oC;         // strides of the block OUTPUT (channels-last, or the [B][T][flat] layout)
    float drop_p;                // 0 disables
    unsigned long long seed;     // dropout seed of this block (block_seed(step seed, block))
    // CUDA-graph steps: the step seed lives in device memory (sedb200_step_state) and is read at run time, so that a
    // captured step draws fresh masks at every replay; null = use `seed`
    const unsigned long long* seed_ptr;
    int block;
    int lgWo, lgH;               // log2 of Wo and H when both are powers of two (every BASELINE config), else -1
};
// window index -> (image, row, window column).  The two divisions by run-time values were a fifth of the instructions
// of the HBM-bound pool kernels (ncu: issue active 64-68 %); shifts when the geometry allows it.
__device__ __forceinline__ void pix_bhw(unsigned pix, const PoolGeom& g, unsigned& b, unsigned& h, unsigned& wo) {
    if (g.lgWo >= 0) {
        wo = pix & ((unsigned)g.Wo - 1u);
        const unsigned t = pix >> g.lgWo;
        h = t & ((unsigned)g.H - 1u);
        b = t >> g.lgH;
    } else {
        const unsigned t = pix / (unsigned)g.Wo;
        wo = pix - t * (unsigned)g.Wo;
        b = t / (unsigned)g.H;
        h = t - b * (unsigned)g.H;
    }
}
__device__ __forceinline__ unsigned long long pool_seed(const PoolGeom& g) {
    return g.seed_ptr ? block_seed(__ldg(g.seed_ptr), g.block) : g.seed;
}

__device__ __forceinline__ void load_dA(const float* __restrict__ da, long oC, float (&g)[4]) {
    if (oC == 1) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(da));
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = __ldg(da + q * oC);
    }
}

}  // namespace sedb200
