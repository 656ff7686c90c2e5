// conv0_lean.cuh -- host entry points of the lean first conv block (conv0_lean.cu); see that file for the method.
#pragma once
#include "crnn_block.cuh"

namespace sedb200 {

// Cin <= 2, 128 | C, pool 2 or 5, and the block output must be channels-last (block 0 is not the last conv block)
bool conv0_lean_ok(int cin, int C, int pool, int n_conv);

int conv0_lean_stats(const float* x, int cin, int C, int H, int W, int batch, const float* w, const float* bias,
                     const float* gamma, const float* beta, float eps, float momentum, float* running, float* stat,
                     double* gram, float* part, cudaStream_t st);

int conv0_lean_forward(const float* x, int cin, int batch, const float* w, const float* bias, const float* stat,
                       const PoolGeom& g, int tensor_cores, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                       unsigned* argw, cudaStream_t st);

// the same forward with a pooling WINDOW per MMA row (conv0_win.cu): max / argmax / ReLU / dropout in registers
bool conv0_win_ok(int cin, int C, int pool, long n_windows);
int conv0_win_forward(const float* x, int cin, int batch, const float* w, const float* bias, const float* stat,
                      const PoolGeom& g, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, unsigned* argw,
                      cudaStream_t st);

long conv0_lean_bwd_part_floats(int cin, int C, int batch, int H);
int conv0_lean_backward(const float* x, int cin, int batch, const unsigned* argw, const float* dA, const PoolGeom& g,
                        const double* gram, const float* w, const float* bias, const float* gamma, const float* stat,
                        float* part, float* dw, float* db, float* dgamma, float* dbeta, cudaStream_t st);

}  // namespace sedb200
