// tc_gemm.cuh -- tcgen05 GEMM (3-term bf16 split) used for the GRU projections and their gradients.
#pragma once
#include "common.cuh"

namespace sedb200 {

bool gemm_tc_supported(int M, int N, int K);
size_t gemm_tc_scratch_bytes(int M, int N, int K, int a_mn, int b_mn, int split_k);

// fp32 [n] -> bf16 hi / lo planes (n % 4 == 0)
int split_planes(const float* x, void* hi, void* lo, long n, cudaStream_t st);
// planes of the [rows][2H] matrix of previous hidden states of both GRU directions
int hprev_planes(const float* out, void* hi, void* lo, long rows, int T, int H, cudaStream_t st);

// D[M][N] (+bias) = sum_k A(m,k) B(n,k); operands as bf16 planes, K-major ([rows][K]) or MN-major ([K][rows])
int gemm_tc(const void* a_hi, const void* a_lo, int a_mn, const void* b_hi, const void* b_lo, int b_mn, int M, int N,
            int K, const float* bias, float* out, long out_ld, int split_k, float* part, cudaStream_t st);

}  // namespace sedb200
