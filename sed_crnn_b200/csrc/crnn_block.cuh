// crnn_block.cuh -- device helpers shared by the conv-block kernels (crnn.cu, conv0_lean.cu): the counter-based
// dropout generator, the bf16 hi / lo plane store, the geometry of a BN + ReLU + max-pool block.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// Activation planes of the plane-native conv blocks: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) -- 22 significand
// bits for the O(1) values BatchNorm produces (residuals below 2^-14 go subnormal: absolute precision 2^-24).  The
// conversions saturate at the fp16 range instead of producing inf.  Storage type of all 16-bit planes is
// `__nv_bfloat16` (just 2 bytes); what the bytes mean is the producer's / consumer's contract.
// {fp16(a), fp16(b)} packed (a in the low half), round-to-nearest, saturating at +-65504: one F2FP instruction
__device__ __forceinline__ unsigned pack_half2_sat(float a, float b) {
    unsigned d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ float2 unpack_half2(unsigned v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ void store_planes4(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long i4,
                                              const float4 v) {
    const unsigned h01 = pack_half2_sat(v.x, v.y), h23 = pack_half2_sat(v.z, v.w);
    const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
    reinterpret_cast<uint2*>(hi)[i4] = make_uint2(h01, h23);
    reinterpret_cast<uint2*>(lo)[i4] = make_uint2(pack_half2_sat(v.x - f01.x, v.y - f01.y), pack_half2_sat(v.z - f23.x, v.w - f23.y));
}
__device__ __forceinline__ float4 load_planes4(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                               long i4) {
    const uint2 hb = __ldg(reinterpret_cast<const uint2*>(hi) + i4);
    const uint2 lb = __ldg(reinterpret_cast<const uint2*>(lo) + i4);
    const __half* hp = reinterpret_cast<const __half*>(&hb);
    const __half* lp = reinterpret_cast<const __half*>(&lb);
    return make_float4(__half2float(hp[0]) + __half2float(lp[0]), __half2float(hp[1]) + __half2float(lp[1]),
                       __half2float(hp[2]) + __half2float(lp[2]), __half2float(hp[3]) + __half2float(lp[3]));
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
};
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
