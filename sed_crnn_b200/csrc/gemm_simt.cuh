// gemm_simt.cuh -- generic fp32 tiled GEMM on the CUDA cores with functor operands.
//
//   C(m, n) = sum_k A(m, k) * B(k, n)          m < M, n < N, k in this CTA's K-slice
//
// A and B are *functors* (implicit operands: im2col gathers, transposed views, two-pointer weight
// sets ...), so one kernel serves conv fwd / dgrad / wgrad, the GRU projections and their gradients
// and the dense layers.  It is the fp32-exact baseline of the CRNN path and the on-device reference
// the tensor-core kernels are validated against.
//
// Tile 128 x 64 x 16, 256 threads, 8 x 4 accumulators per thread, register-staged double buffering.
// Split-K: blockIdx.z selects a K-slice; the epilogue receives the slice index and writes a partial
// that a second, fixed-order pass reduces (deterministic -- no float atomics anywhere).
#pragma once
#include "common.cuh"

namespace sedb200 {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

// Operand functor contract:
//   static constexpr bool kContigK;          // true: consecutive k are adjacent in memory
//   __device__ float operator()(int row, int k) const;   // A: row = m ; B: row = n
// Epilogue contract:
//   __device__ void operator()(int m, int n, float acc, int kslice) const;

template <class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(GTHREADS)
gemm_simt_kernel(int M, int N, int K, int k_slice, AOp A, BOp B, Epi epi) {
    pdl_wait();
    __shared__ __align__(16) float As[2][GBK][GBM + 4];
    __shared__ __align__(16) float Bs[2][GBK][GBN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int kbeg = blockIdx.z * k_slice;
    const int kend = min(K, kbeg + k_slice);
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 thread grid: 4 cols x 8 rows each

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    constexpr int A_PER = GBM * GBK / GTHREADS;      // 8
    constexpr int B_PER = GBN * GBK / GTHREADS;      // 4
    float ra[A_PER], rb[B_PER];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int mm, kk;
            if (AOp::kContigK) { kk = tid % GBK; mm = tid / GBK + i * (GTHREADS / GBK); }
            else               { mm = tid % GBM; kk = tid / GBM + i * (GTHREADS / GBM); }
            const int m = m0 + mm, k = k0 + kk;
            ra[i] = (m < M && k < kend) ? A(m, k) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int nn, kk;
            if (BOp::kContigK) { kk = tid % GBK; nn = tid / GBK + i * (GTHREADS / GBK); }
            else               { nn = tid % GBN; kk = tid / GBN + i * (GTHREADS / GBN); }
            const int n = n0 + nn, k = k0 + kk;
            rb[i] = (n < N && k < kend) ? B(n, k) : 0.0f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int mm, kk;
            if (AOp::kContigK) { kk = tid % GBK; mm = tid / GBK + i * (GTHREADS / GBK); }
            else               { mm = tid % GBM; kk = tid / GBM + i * (GTHREADS / GBM); }
            As[buf][kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int nn, kk;
            if (BOp::kContigK) { kk = tid % GBK; nn = tid / GBK + i * (GTHREADS / GBK); }
            else               { nn = tid % GBN; kk = tid / GBN + i * (GTHREADS / GBN); }
            Bs[buf][kk][nn] = rb[i];
        }
    };

    if (kbeg < kend) {
        fetch(kbeg);
        stash(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = kbeg; k0 < kend; k0 += GBK) {
            const bool more = k0 + GBK < kend;
            if (more) fetch(k0 + GBK);
#pragma unroll
            for (int kk = 0; kk < GBK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (more) {
                stash(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epi(m, n, acc[i][j], (int)blockIdx.z);
        }
    }
}

// Launch helper.  `splits` >= 1 K-slices (slice length rounded up to a multiple of GBK).
template <class AOp, class BOp, class Epi>
inline int gemm_simt(int M, int N, int K, int splits, AOp A, BOp B, Epi epi, cudaStream_t st) {
    if (M <= 0 || N <= 0) return SEDB200_OK;
    if (splits < 1) splits = 1;
    int k_slice = (K + splits - 1) / splits;
    k_slice = ((k_slice + GBK - 1) / GBK) * GBK;
    if (k_slice < GBK) k_slice = GBK;
    splits = K > 0 ? (K + k_slice - 1) / k_slice : 1;
    dim3 grid((M + GBM - 1) / GBM, (N + GBN - 1) / GBN, splits);
    launch_k(gemm_simt_kernel<AOp, BOp, Epi>, grid, GTHREADS, 0, st, M, N, K, k_slice, A, B, epi);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// number of K-slices the helper will actually launch for a requested split count
inline int gemm_simt_splits(int K, int splits) {
    if (splits < 1) splits = 1;
    int k_slice = (K + splits - 1) / splits;
    k_slice = ((k_slice + GBK - 1) / GBK) * GBK;
    if (k_slice < GBK) k_slice = GBK;
    return K > 0 ? (K + k_slice - 1) / k_slice : 1;
}

// ---- fixed-order reduction of split-K partials: out[i] = sum_z part[z][i]  (+ optional accumulate)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, long n,
                                       int splits);
int reduce_partials(const float* part, float* out, long n, int splits, cudaStream_t st);

// ---- deterministic column sums: out[c] = sum_r X[r][c], X row-major [rows][cols]
//      `scratch` needs colsum_scratch_floats(rows, cols) floats.
long colsum_scratch_floats(long rows, int cols);
int colsum(const float* X, long rows, int cols, float* out, float* scratch, cudaStream_t st);
//      first stage only: part[nblk][2][cols] = per-block {sum x, sum x^2}; nblk = colsum_blocks(rows)
int colsum_blocks(long rows);
int colsum_partials(const float* X, long rows, int cols, float* part, int* nblk_out, cudaStream_t st);

}  // namespace sedb200
