// conv_small.cuh -- direct fp32 3x3 convolutions for small channel counts (<= 64), see conv_small.cu.
#pragma once
#include "common.cuh"

namespace sedb200 {

// forward (dgrad = 0): out[B][H][W][N] = conv3x3(in, wgt[N][K][3][3]) + bias;  in is read through strides
//   (element b*sB + h*sH + w*sW + k*sC), so NCHW user input and channels-last activations both work.
// data gradient (dgrad = 1): in = dY [B][H][W][K] (K = Cout of the forward conv), wgt = the forward weight
//   [K][N][3][3], out = dX [B][H][W][N].
bool conv_small_supported(int K, int N, int W);
int  conv_small_forward(const float* in, long sB, long sH, long sW, long sC, int K, int B, int H, int W,
                        const float* wgt, const float* bias, int N, int dgrad, float* out, cudaStream_t st);

// weight gradient: dw[N][K][3][3] from dy [B][H][W][N] and the (strided) conv input with K channels.
// `part` needs conv_small_wgrad_part_floats() floats.
bool   conv_small_wgrad_supported(int K, int N, int W);
size_t conv_small_wgrad_part_floats(int K, int N, int B, int H);
int    conv_small_wgrad(const float* dy, const float* in, long sB, long sH, long sW, long sC, int K, int N, int B,
                        int H, int W, float* part, float* dw, cudaStream_t st);

}  // namespace sedb200
