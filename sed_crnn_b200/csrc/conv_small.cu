// conv_small.cu -- direct fp32 3x3 convolutions for SMALL channel counts (the reference's shipped configuration,
// train_constants.py: CONV_DEPTH = 16; crnn_lightning.py:46-51).
//
// With 16 or 32 channels a conv block is far too small for the tensor cores' 128-wide tiles and far too gather-heavy
// for the generic functor GEMM (every operand element costs an im2col index computation with three integer
// divisions).  These kernels keep one band of 8 image rows (plus halo) of the input in shared memory and
//   forward / data gradient: thread <-> 4 neighbouring pixels x 8 output channels (32 accumulators); per (input
//       channel, kernel row) the 6 input columns the 4 pixels touch are read once and the 3 x 8 weights as two
//       broadcast LDS.128 each -- 96 FMAs per 9 shared-memory loads;
//   weight gradient: thread <-> one (output channel, input channel) pair (or, for a 1-channel input, one (output
//       channel, band row)) with its 9 taps in registers for the whole band; per-block partials, fixed-order
//       reduction afterwards (no float atomics).
// The data gradient is the forward kernel run on dY with the weights flipped and transposed while they are staged
// into shared memory, so no separate weight-transform kernel exists.
#include "conv_small.cuh"

#include <algorithm>
#include <cstdlib>

namespace sedb200 {
namespace {

constexpr int kRows = 8;               // image rows per block

__host__ __device__ inline int in_pitch(int W) { return W + 4; }     // even pitch: 8 B aligned column pairs
inline size_t direct_smem(int K, int N, int W) {
    return ((size_t)K * (kRows + 2) * in_pitch(W) + (size_t)9 * K * N) * 4;
}

// in : strided input [b][k][h][w] -> element b*sB + h*sH + w*sW + k*sC   (K input channels)
// out: channels-last [B][H][W][N]
// wgt: PyTorch layout.  dgrad == 0: wgt[n][k][tap] (N = Cout, K = Cin);  dgrad == 1: wgt[k][n][tap] (K = Cout of the
//      forward conv, N = its Cin) and the taps are flipped.
__global__ void __launch_bounds__(512)
conv_small_kernel(const float* __restrict__ in, long sB, long sH, long sW, long sC, int K, int H, int W,
                  const float* __restrict__ wgt, const float* __restrict__ bias, int N, int dgrad,
                  float* __restrict__ out) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int Wp = in_pitch(W);
    float* xs = sm;                                  // [K][kRows+2][Wp]   column w of the image sits at index w + 1
    float* ws = xs + (size_t)K * (kRows + 2) * Wp;   // [9][K][N]
    const int b = blockIdx.y, h0 = blockIdx.x * kRows;
    for (int i = threadIdx.x; i < K * (kRows + 2) * Wp; i += blockDim.x) {
        const int cc = i % Wp, rr = (i / Wp) % (kRows + 2), k = i / (Wp * (kRows + 2));
        const int hh = h0 - 1 + rr, ww = cc - 1;
        xs[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(in + (long)b * sB + (long)hh * sH + (long)ww * sW + (long)k * sC) : 0.0f;
    }
    for (int i = threadIdx.x; i < 9 * K * N; i += blockDim.x) {
        const int n = i % N, k = (i / N) % K, tap = i / (N * K);
        ws[i] = dgrad ? __ldg(wgt + ((long)k * N + n) * 9 + (8 - tap)) : __ldg(wgt + ((long)n * K + k) * 9 + tap);
    }
    __syncthreads();
    // thread -> (band row, pixel quad, channel group of 8)
    const int NG = N >> 3, Q = W >> 2;
    const int g = threadIdx.x % NG, q = (threadIdx.x / NG) % Q, row = threadIdx.x / (NG * Q);
    const int h = h0 + row;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][c] = bias ? __ldg(bias + 8 * g + c) : 0.0f;
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* xr = xs + ((size_t)k * (kRows + 2) + row + r) * Wp + 4 * q;       // columns 4q-1 .. 4q+4
            float xc[6];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const float2 v = *reinterpret_cast<const float2*>(xr + 2 * t);
                xc[2 * t] = v.x; xc[2 * t + 1] = v.y;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float4* wp = reinterpret_cast<const float4*>(ws + ((size_t)(r * 3 + s) * K + k) * N + 8 * g);
                const float4 w0 = wp[0], w1 = wp[1];
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(xc[p + s], wv[c], acc[p][c]);
            }
        }
    }
    if (h < H) {
        float* o = out + (((long)b * H + h) * W + 4 * q) * N + 8 * g;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            *reinterpret_cast<float4*>(o + (long)p * N) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
            *reinterpret_cast<float4*>(o + (long)p * N + 4) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
        }
    }
}

// part[blk][n][k][tap] = sum over the block's band of dy[p][n] * in[p shifted by tap][k]
// threads = N * K * RS: (n, k) pair x row slice; RS row slices share the band's rows round-robin and are folded in
// shared memory at the end (fixed order).
__global__ void __launch_bounds__(1024)
conv_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ in, long sB, long sH, long sW, long sC,
                        int K, int N, int H, int W, int RS, float* __restrict__ part) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int Wp = W + 2;
    float* xs = sm;                                  // [kRows+2][Wp][K]   (k fastest: lanes of a warp read consecutive words)
    float* ds = xs + (size_t)(kRows + 2) * Wp * K;   // [kRows][W][N]
    float* red = ds + (size_t)kRows * W * N;         // [RS][N*K*9]  (only when RS > 1)
    const int b = blockIdx.y, h0 = blockIdx.x * kRows;
    for (int i = threadIdx.x; i < (kRows + 2) * Wp * K; i += blockDim.x) {
        const int k = i % K, cc = (i / K) % Wp, rr = i / (K * Wp);
        const int hh = h0 - 1 + rr, ww = cc - 1;
        xs[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(in + (long)b * sB + (long)hh * sH + (long)ww * sW + (long)k * sC) : 0.0f;
    }
    for (int i = threadIdx.x; i < kRows * W * N; i += blockDim.x) {
        const int rr = i / (W * N);
        ds[i] = (h0 + rr) < H ? __ldg(dy + ((long)b * H + h0) * W * N + i) : 0.0f;
    }
    __syncthreads();
    const int k = threadIdx.x % K, n = (threadIdx.x / K) % N, rs = threadIdx.x / (K * N);
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.0f;
    for (int row = rs; row < kRows; row += RS) {
        for (int w = 0; w < W; ++w) {
            const float d = ds[((size_t)row * W + w) * N + n];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    acc[r * 3 + s] = fmaf(d, xs[((size_t)(row + r) * Wp + w + s) * K + k], acc[r * 3 + s]);
        }
    }
    const long blk = (long)blockIdx.y * gridDim.x + blockIdx.x;
    float* pb = part + blk * (long)N * K * 9 + ((long)n * K + k) * 9;
    if (RS == 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t) pb[t] = acc[t];
        return;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) red[((size_t)rs * N * K + (size_t)n * K + k) * 9 + t] = acc[t];
    __syncthreads();
    if (rs == 0) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float a = 0.0f;
            for (int z = 0; z < RS; ++z) a += red[((size_t)z * N * K + (size_t)n * K + k) * 9 + t];
            pb[t] = a;
        }
    }
}

// dW[i] = sum_blk part[blk][i]: one warp per output, fixed order
__global__ void conv_small_wgrad_reduce_kernel(const float* __restrict__ part, int nblk, int n_out, float* __restrict__ dw) {
    pdl_wait();
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= n_out) return;
    double a = 0.0;
    for (int z = lane; z < nblk; z += 32) a += (double)__ldg(part + (long)z * n_out + o);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) dw[o] = (float)a;
}

inline int wgrad_row_slices(int K, int N) {
    int rs = 256 / (K * N);
    rs = std::max(1, std::min(rs, kRows));
    while (kRows % rs) --rs;
    return rs;
}
inline size_t wgrad_smem(int K, int N, int W, int RS) {
    return ((size_t)(kRows + 2) * (W + 2) * K + (size_t)kRows * W * N + (RS > 1 ? (size_t)RS * N * K * 9 : 0)) * 4;
}

}  // namespace

bool conv_small_supported(int K, int N, int W) {
    if (getenv("SEDB200_NO_CONV_SMALL")) return false;
    if (N % 8 || W % 4 || N > 64 || K > 64 || K < 1) return false;
    const long threads = (long)kRows * (W / 4) * (N / 8);
    return threads >= 32 && threads <= 512 && direct_smem(K, N, W) <= 200 * 1024;
}

int conv_small_forward(const float* in, long sB, long sH, long sW, long sC, int K, int B, int H, int W, const float* wgt,
                       const float* bias, int N, int dgrad, float* out, cudaStream_t st) {
    SED_REQUIRE(conv_small_supported(K, N, W), SEDB200_ESHAPE, "conv_small: K=%d N=%d W=%d unsupported", K, N, W);
    const size_t smem = direct_smem(K, N, W);
    int rc = ensure_dyn_smem((const void*)conv_small_kernel, 200 * 1024);
    if (rc) return rc;
    const dim3 grid((H + kRows - 1) / kRows, B);
    launch_k(conv_small_kernel, grid, kRows * (W / 4) * (N / 8), smem, st, in, sB, sH, sW, sC, K, H, W, wgt, bias, N, dgrad, out);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

bool conv_small_wgrad_supported(int K, int N, int W) {
    if (getenv("SEDB200_NO_CONV_SMALL_WGRAD") || getenv("SEDB200_NO_CONV_SMALL")) return false;
    if (K * N > 1024 || K < 1 || N < 1) return false;
    return wgrad_smem(K, N, W, wgrad_row_slices(K, N)) <= 200 * 1024;
}

size_t conv_small_wgrad_part_floats(int K, int N, int B, int H) {
    return (size_t)B * ((H + kRows - 1) / kRows) * N * K * 9;
}

int conv_small_wgrad(const float* dy, const float* in, long sB, long sH, long sW, long sC, int K, int N, int B, int H,
                     int W, float* part, float* dw, cudaStream_t st) {
    SED_REQUIRE(conv_small_wgrad_supported(K, N, W), SEDB200_ESHAPE, "conv_small_wgrad: K=%d N=%d W=%d unsupported", K, N, W);
    const int RS = wgrad_row_slices(K, N);
    int rc = ensure_dyn_smem((const void*)conv_small_wgrad_kernel, 200 * 1024);
    if (rc) return rc;
    const dim3 grid((H + kRows - 1) / kRows, B);
    launch_k(conv_small_wgrad_kernel, grid, K * N * RS, wgrad_smem(K, N, W, RS), st, dy, in, sB, sH, sW, sC, K, N, H, W, RS, part);
    SED_POST_LAUNCH();
    const int n_out = N * K * 9, nblk = (int)(grid.x * grid.y);
    launch_k(conv_small_wgrad_reduce_kernel, (n_out * 32 + 255) / 256, 256, 0, st, part, nblk, n_out, dw);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // namespace sedb200

using namespace sedb200;

extern "C" {

int sedb200_conv3x3_small_supported(int K, int N, int W, int wgrad) {
    return (wgrad ? conv_small_wgrad_supported(K, N, W) : conv_small_supported(K, N, W)) ? 1 : 0;
}

int sedb200_conv3x3_small(const float* in_dev, long sB, long sH, long sW, long sC, int K, int B, int H, int W,
                          const float* weight_dev, const float* bias_dev, int N, int dgrad, float* out_dev,
                          void* stream) {
    SED_REQUIRE(in_dev && weight_dev && out_dev && B >= 1 && H >= 1, SEDB200_EINVAL, "conv3x3_small: bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    return conv_small_forward(in_dev, sB, sH, sW, sC, K, B, H, W, weight_dev, bias_dev, N, dgrad, out_dev, as_stream(stream));
}

size_t sedb200_conv3x3_small_wgrad_scratch_bytes(int K, int N, int B, int H) {
    return conv_small_wgrad_part_floats(K, N, B, H) * 4;
}

int sedb200_conv3x3_small_wgrad(const float* dy_dev, const float* in_dev, long sB, long sH, long sW, long sC, int K,
                                int N, int B, int H, int W, float* dw_dev, void* scratch_dev, size_t scratch_bytes,
                                void* stream) {
    SED_REQUIRE(dy_dev && in_dev && dw_dev && scratch_dev && B >= 1 && H >= 1, SEDB200_EINVAL, "conv3x3_small_wgrad: bad argument");
    SED_REQUIRE(scratch_bytes >= sedb200_conv3x3_small_wgrad_scratch_bytes(K, N, B, H), SEDB200_EWORKSPACE,
                "conv3x3_small_wgrad: scratch too small");
    int rc = require_sm100();
    if (rc) return rc;
    return conv_small_wgrad(dy_dev, in_dev, sB, sH, sW, sC, K, N, B, H, W, reinterpret_cast<float*>(scratch_dev), dw_dev,
                            as_stream(stream));
}

}  // extern "C"
