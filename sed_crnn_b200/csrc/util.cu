// util.cu -- deterministic reductions shared by the CRNN kernels (no float atomics anywhere).
#include "gemm_simt.cuh"

#include <algorithm>

namespace sedb200 {

__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, long n,
                                       int splits) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < splits; ++z) s += part[(long)z * n + i];
        out[i] = s;
    }
}

int reduce_partials(const float* part, float* out, long n, int splits, cudaStream_t st) {
    if (n <= 0) return SEDB200_OK;
    const int blocks = (int)std::min<long>((n + 255) / 256, 1184);
    reduce_partials_kernel<<<blocks, 256, 0, st>>>(part, out, n, splits);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

// ---------------------------------------------------------------- column sums (sum and sum of squares)
namespace {
constexpr int kColBlocksMax = 592;

__host__ __device__ inline int col_width(int cols) {
    int w = 1;
    while (w < cols && w < 256) w <<= 1;
    return w;
}

// part layout: [nblk][2][cols]  (0: sum x, 1: sum x^2)
__global__ void __launch_bounds__(256)
colsum2_kernel(const float* __restrict__ X, long rows, int cols, long rows_per_blk, float* __restrict__ part) {
    __shared__ float s1[256], s2[256];
    const int CW = col_width(cols), RY = 256 / CW;
    const int cx = threadIdx.x % CW, ry = threadIdx.x / CW;
    const long r0 = (long)blockIdx.x * rows_per_blk;
    const long r1 = min(rows, r0 + rows_per_blk);
    for (int c0 = 0; c0 < cols; c0 += CW) {
        const int c = c0 + cx;
        float a = 0.0f, b = 0.0f;
        if (c < cols)
            for (long r = r0 + ry; r < r1; r += RY) {
                const float x = __ldg(X + r * cols + c);
                a += x;
                b = fmaf(x, x, b);
            }
        s1[threadIdx.x] = a;
        s2[threadIdx.x] = b;
        __syncthreads();
        if (ry == 0 && c < cols) {
            float ta = 0.0f, tb = 0.0f;
            for (int y = 0; y < RY; ++y) {
                ta += s1[y * CW + cx];
                tb += s2[y * CW + cx];
            }
            part[((long)blockIdx.x * 2 + 0) * cols + c] = ta;
            part[((long)blockIdx.x * 2 + 1) * cols + c] = tb;
        }
        __syncthreads();
    }
}

// one warp per column: fixed lane assignment + fixed shuffle tree (deterministic)
__global__ void colsum_final_kernel(const float* __restrict__ part, int nblk, int cols, float* __restrict__ out,
                                    float* __restrict__ out_sq) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= cols) return;
    double a = 0.0, b = 0.0;
    for (int k = lane; k < nblk; k += 32) {
        a += (double)__ldg(part + ((long)k * 2 + 0) * cols + c);
        b += (double)__ldg(part + ((long)k * 2 + 1) * cols + c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        if (out) out[c] = (float)a;
        if (out_sq) out_sq[c] = (float)b;
    }
}
}  // namespace

int colsum_blocks(long rows) {
    long nb = (rows + 63) / 64;
    if (nb < 1) nb = 1;
    if (nb > kColBlocksMax) nb = kColBlocksMax;
    return (int)nb;
}

long colsum_scratch_floats(long rows, int cols) { return (long)colsum_blocks(rows) * 2 * cols; }

int colsum_partials(const float* X, long rows, int cols, float* part, int* nblk_out, cudaStream_t st) {
    const int nb = colsum_blocks(rows);
    const long rpb = (rows + nb - 1) / nb;
    colsum2_kernel<<<nb, 256, 0, st>>>(X, rows, cols, rpb, part);
    SED_POST_LAUNCH();
    *nblk_out = nb;
    return SEDB200_OK;
}

int colsum(const float* X, long rows, int cols, float* out, float* scratch, cudaStream_t st) {
    int nb = 0;
    int rc = colsum_partials(X, rows, cols, scratch, &nb, st);
    if (rc) return rc;
    colsum_final_kernel<<<(cols * 32 + 255) / 256, 256, 0, st>>>(scratch, nb, cols, out, nullptr);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace sedb200
