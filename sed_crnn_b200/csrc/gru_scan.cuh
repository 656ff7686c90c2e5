// gru_scan.cuh -- recurrent part of the bidirectional GRU (nn.GRU semantics, gate order r,z,n).
#pragma once
#include "common.cuh"

namespace sedb200 {

// gi    [B][T][2][3H]  input projections incl. b_ih (direction 0 = forward, 1 = reverse)
// whh   [2][3H][H], bhh [2][3H]
// out   [B][T][2H]     layer output (forward half | reverse half)
// gates [B][T][2][4H]  r, z, n and q = W_hn h_{t-1} + b_hn, saved for the backward scan
int gru_scan_forward(const float* gi, const float* whh, const float* bhh, float* out, float* gates,
                     int B, int T, int H, cudaStream_t st);

// dout  [B][T][2H] gradient w.r.t. the layer output
// dgi   [B][T][2][3H] gradient w.r.t. gi (= input-side pre-activations)
// dgh   [B][T][2][3H] gradient w.r.t. W_hh h_{t-1} + b_hh
// When gru_scan_fused_param_grads(H) the scan also writes per-batch-row partial sums that the caller
// reduces over B in a fixed order:
//   part_w                 unused (kept for ABI stability; dW_hh comes from a GEMM over dgh and h_{t-1})
//   part_b [B][2][2][3H]   contribution of row b to db_ih ([.][0]) and db_hh ([.][1])
// otherwise part_w / part_b are untouched and the caller derives those gradients from dgi / dgh.
bool gru_scan_fused_param_grads(int H);
int gru_scan_backward(const float* dout, const float* out, const float* gates, const float* whh,
                      float* dgi, float* dgh, float* part_w, float* part_b, int B, int T, int H,
                      cudaStream_t st);

}  // namespace sedb200
