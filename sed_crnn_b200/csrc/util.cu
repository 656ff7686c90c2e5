// util.cu -- deterministic reductions shared by the CRNN kernels (no float atomics anywhere).
#include "gemm_simt.cuh"

#include <algorithm>

namespace sedb200 {

__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, long n,
                                       int splits) {
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        out[i] = ordered_sum<8, float>(part + i, n, splits);
    }
}

int reduce_partials(const float* part, float* out, long n, int splits, cudaStream_t st) {
    if (n <= 0) return SEDB200_OK;
    const int blocks = (int)std::min<long>((n + 255) / 256, 1184);
    launch_k(reduce_partials_kernel, blocks, 256, 0, st, part, out, n, splits);
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
    pdl_wait();
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
    pdl_wait();
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
    launch_k(colsum2_kernel, nb, 256, 0, st, X, rows, cols, rpb, part);
    SED_POST_LAUNCH();
    *nblk_out = nb;
    return SEDB200_OK;
}

int colsum(const float* X, long rows, int cols, float* out, float* scratch, cudaStream_t st) {
    int nb = 0;
    int rc = colsum_partials(X, rows, cols, scratch, &nb, st);
    if (rc) return rc;
    launch_k(colsum_final_kernel, (cols * 32 + 255) / 256, 256, 0, st, scratch, nb, cols, out, nullptr);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}


// ---------------------------------------------------------------- per-bin standardisation (feature.py:127-129)
namespace {
// one warp per column: mean, population variance, scale (= sqrt(var), 1.0 for constant columns) in float64
__global__ void standardize_final_kernel(const float* __restrict__ part, int nblk, int cols, long rows,
                                         double* __restrict__ mean, double* __restrict__ var,
                                         double* __restrict__ scale) {
    pdl_wait();
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
        const double m = a / (double)rows;
        double v = b / (double)rows - m * m;
        if (v < 0.0) v = 0.0;
        mean[c] = m;
        var[c] = v;
        // sklearn _handle_zeros_in_scale: (near-)constant features keep scale 1
        scale[c] = (v <= 10.0 * 2.220446049250313e-16 * (m * m + 1e-300) || v == 0.0) ? 1.0 : sqrt(v);
    }
}
__global__ void __launch_bounds__(256)
standardize_apply_kernel(const float* __restrict__ X, long n, int cols, const double* __restrict__ mean,
                         const double* __restrict__ scale, float* __restrict__ out) {
    pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        out[i] = (float)((float)((double)__ldg(X + i) - mean[c]) / scale[c]);      // two roundings, like numpy in-place ops
    }
}
}  // namespace

}  // namespace sedb200

using namespace sedb200;

extern "C" {

size_t sedb200_standardize_scratch_bytes(long rows, int cols) { return (size_t)colsum_scratch_floats(rows, cols) * 4; }

int sedb200_standardize_fit(const float* x_dev, long rows, int cols, double* mean_dev, double* var_dev,
                            double* scale_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(x_dev && mean_dev && var_dev && scale_dev && scratch_dev, SEDB200_EINVAL, "standardize_fit: null buffer");
    SED_REQUIRE(rows >= 1 && cols >= 1, SEDB200_EINVAL, "standardize_fit: rows=%ld cols=%d", rows, cols);
    SED_REQUIRE(scratch_bytes >= sedb200_standardize_scratch_bytes(rows, cols), SEDB200_EWORKSPACE, "standardize_fit: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    int nb = 0;
    rc = colsum_partials(x_dev, rows, cols, reinterpret_cast<float*>(scratch_dev), &nb, st);
    if (rc) return rc;
    launch_k(standardize_final_kernel, (cols * 32 + 255) / 256, 256, 0, st, reinterpret_cast<float*>(scratch_dev), nb, cols, rows,
                                                                    mean_dev, var_dev, scale_dev);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_standardize_apply(const float* x_dev, long rows, int cols, const double* mean_dev, const double* scale_dev,
                              float* out_dev, void* stream) {
    SED_REQUIRE(x_dev && mean_dev && scale_dev && out_dev, SEDB200_EINVAL, "standardize_apply: null buffer");
    SED_REQUIRE(rows >= 0 && cols >= 1, SEDB200_EINVAL, "standardize_apply: rows=%ld cols=%d", rows, cols);
    if (rows == 0) return SEDB200_OK;
    int rc = require_sm100();
    if (rc) return rc;
    const long n = rows * cols;
    launch_k(standardize_apply_kernel, (int)std::min<long>((n + 255) / 256, 148L * 16), 256, 0, as_stream(stream),
        x_dev, n, cols, mean_dev, scale_dev, out_dev);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // extern "C"
