// gru_scan.cuh -- recurrent part of the bidirectional GRU (nn.GRU semantics, gate order r,z,n).
#pragma once
#include "common.cuh"

namespace sedb200 {

// gi    [B][T][2][3H]  input projections incl. b_ih (direction 0 = forward, 1 = reverse)
// whh   [2][3H][H], bhh [2][3H]
// out   [B][T][2H]     layer output (forward half | reverse half)
// gates [B][T][2][4H]  r, z, n and q = W_hn h_{t-1} + b_hn, saved for the backward scan
// hprev_hi / hprev_lo (optional, only when gru_scan_emits_planes(H)): bf16 hi/lo planes of h_{t-1} as one
//       [B*T][2H] matrix -- columns [0,H) = out[b][t-1][0:H] (0 at t = 0), columns [H,2H) = out[b][t+1][H:2H]
//       (0 at t = T-1) -- the B operand of the dW_hh tensor-core GEMM, written by the scan itself
bool gru_scan_emits_planes(int H);
int gru_scan_forward(const float* gi, const float* whh, const float* bhh, float* out, float* gates,
                     int B, int T, int H, cudaStream_t st, void* hprev_hi = nullptr, void* hprev_lo = nullptr);

// dout  [B][T][2H] gradient w.r.t. the layer output
// dgi   [B][T][2][3H] gradient w.r.t. gi (= input-side pre-activations)
// dgh   [B][T][2][3H] gradient w.r.t. W_hh h_{t-1} + b_hh
// When gru_scan_fused_param_grads(H) the scan also writes per-batch-row partial sums that the caller
// reduces over B in a fixed order:
//   part_w                 unused (kept for ABI stability; dW_hh comes from a GEMM over dgh and h_{t-1})
//   part_b [B][2][2][3H]   contribution of row b to db_ih ([.][0]) and db_hh ([.][1])
// otherwise part_w / part_b are untouched and the caller derives those gradients from dgi / dgh.
bool gru_scan_fused_param_grads(int H);
// planes (optional, only when gru_scan_emits_planes(H)): {dgi_hi, dgi_lo, dgh_hi, dgh_lo}; when given, dgi / dgh are
// written ONLY as bf16 hi/lo planes (the operands of the tensor-core GEMMs that consume them) and the fp32 arrays
// are left untouched.
int gru_scan_backward(const float* dout, const float* out, const float* gates, const float* whh,
                      float* dgi, float* dgh, float* part_w, float* part_b, int B, int T, int H,
                      cudaStream_t st, void* const* planes = nullptr);

}  // namespace sedb200
