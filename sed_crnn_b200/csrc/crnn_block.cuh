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
// One 64-bit hash (splitmix64 finaliser) per group of FOUR consecutive elements; element q of the group keeps its
// value iff the q-th 16-bit field of the hash is >= p * 65536.  Forward and backward kernels all go through
// dropout_keep4(seed, group index), so they agree on the mask without storing it.
__device__ __forceinline__ unsigned long long hash64(unsigned long long seed, unsigned long long idx) {
    unsigned long long x = seed + idx * 0x9E3779B97F4A7C15ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
struct Keep4 { bool k[4]; };
__device__ __forceinline__ Keep4 dropout_keep4(unsigned long long seed, unsigned long long group, float p) {
    const unsigned long long x = hash64(seed, group);
    const unsigned thr = (unsigned)(p * 65536.0f);
    Keep4 r;
#pragma unroll
    for (int q = 0; q < 4; ++q) r.k[q] = (unsigned)((x >> (16 * q)) & 0xFFFFu) >= thr;
    return r;
}
__host__ __device__ inline unsigned long long block_seed(unsigned long long seed, int block) {
    return seed * 0x2545F4914F6CDD1Dull + (unsigned long long)(block + 1) * 0xD6E8FEB86659FD93ull;
}

// Activation planes of the plane-native conv blocks: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) -- 22 significand
// bits for the O(1) values BatchNorm produces (residuals below 2^-14 go subnormal: absolute precision 2^-24).  The
// conversions saturate at the fp16 range instead of producing inf.  Storage type of all 16-bit planes is
// `__nv_bfloat16` (just 2 bytes); what the bytes mean is the producer's / consumer's contract.
__device__ __forceinline__ __half sat_half(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)); }
__device__ __forceinline__ void store_planes4(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long i4,
                                              const float4 v) {
    __half h[4], l[4];
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        h[q] = sat_half(f[q]);
        l[q] = sat_half(f[q] - __half2float(h[q]));
    }
    reinterpret_cast<uint2*>(hi)[i4] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i4] = *reinterpret_cast<uint2*>(l);
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
    __half h[4] = {sat_half(v.x * scale), sat_half(v.y * scale), sat_half(v.z * scale), sat_half(v.w * scale)};
    reinterpret_cast<uint2*>(hi)[i4] = *reinterpret_cast<uint2*>(h);
}

struct PoolGeom {
    int H, W, Wo, C, p;
    long oB, oH, oW, oC;         // strides of the block OUTPUT (channels-last, or the [B][T][flat] layout)
    float drop_p;                // 0 disables
    unsigned long long seed;
};

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
