// tc_conv.cuh -- tensor-core (tcgen05) 3x3 convolution entry points used by the CRNN orchestration.
#pragma once
#include "common.cuh"

namespace sedb200 {

// shapes the tcgen05 path covers: K channels % 64 == 0, N channels % 128 == 0, W a divisor of 128
bool conv_tc_supported(int H, int W, int Kc, int Nc);
size_t conv_tc_scratch_bytes(int B, int H, int W, int Kc, int Nc);

// dgrad == 0: out[B,H,W,Cout] = conv3x3(in[B,H,W,Cin], w[Cout][Cin][3][3]) + bias      (channels-last fp32)
// dgrad == 1: out[B,H,W,Cin]  = conv3x3_transposed(in = dY[B,H,W,Cout], w)             (bias ignored: pass null)
int conv_tc_forward(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin,
                    int Cout, int dgrad, void* scratch, size_t scratch_bytes, cudaStream_t st);

// dw[Cout][Cin][3][3] = sum over pixels of dy[B,H,W,Cout] (x) shifted in[B,H,W,Cin]
// shapes: Cin % 128 == 0, Cout % 128 == 0, W a divisor of 32
bool wgrad_tc_supported(int H, int W, int Cin, int Cout);
size_t wgrad_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout);
int wgrad_tc(const float* dy, const float* in, float* dw, int B, int H, int W, int Cin, int Cout, void* scratch,
             size_t scratch_bytes, cudaStream_t st);

// what the 16-bit operand planes hold: bf16 hi / lo (the fp32 entry points split on the fly) or fp16 (the plane-native
// CRNN flow: fp16 hi / lo activations and weights, one scaled fp16 plane for gradients -- crnn_block.cuh)
constexpr int kPlaneBF16 = 0, kPlaneF16 = 1;
// terms == 2 (forward of the plane-native blocks): an fp16 hi*hi pass into one accumulator plus ONE fp8 pass over the
// combined correction planes (crnn_block.cuh: c8) into a second accumulator that the epilogue adds with the scale
// out2_scale[0] = 2^-(12 + b), 2^b being the weight tensor's fp8 scale (conv_tc_weight_planes writes both)

// ---- plane-native variants: the activation / gradient tensors already exist as 16-bit hi / lo planes
size_t conv_tc_weight_scratch_bytes(int Cin, int Cout);
// stats (optional): per-M-tile partial BatchNorm sums [conv_tc_stat_tiles()][2][Cout] of the conv output (+bias)
int conv_tc_stat_tiles(int B, int H, int W);
int conv_tc_planes(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* out, float* stats,
                   int B, int H, int W, int Cin, int Cout, int dgrad, void* wscratch, cudaStream_t st);
int conv_tc_weight_planes(const float* w, int Cin, int Cout, int dgrad, void* wplanes, cudaStream_t st,
                          int fmt = kPlaneBF16, float* scale2 = nullptr);
// scale2 (device, 2 floats, fp16 forward layout only): receives {2^b, 2^-(12 + b)}; the second plane then holds the
// combined e4m3 correction operand [tap][N][K/64][lo8 x 64 | hi8 x 64] instead of the fp16 lo plane
// terms: 3 = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi (fp32-grade); 1 = hi planes only (the lo pointers are not read)
int conv_tc_planes_w(const void* a_hi, const void* a_lo, const void* wplanes, const float* bias, float* out, float* stats,
                     int B, int H, int W, int Cin, int Cout, int dgrad, cudaStream_t st, int terms = 3,
                     int fmt = kPlaneBF16, const float* out_scale = nullptr, const float* out2_scale = nullptr);
// out_scale (device pointer, optional): every accumulator is multiplied by out_scale[0] before bias / store
size_t wgrad_tc_part_bytes(int Cin, int Cout);
int wgrad_tc_planes(const void* y_hi, const void* y_lo, const void* x_hi, const void* x_lo, float* dw, int B, int H,
                    int W, int Cin, int Cout, float* part, cudaStream_t st, int terms = 3, int fmt = kPlaneBF16,
                    const float* out_scale = nullptr, int max_stages = 0);
// max_stages (optional): cap of the pipeline depth = shared memory per CTA (32 KB per single-pass stage), for callers
// that run the kernel beside another one on the same SMs

}  // namespace sedb200
