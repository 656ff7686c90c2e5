// crnn_plan.cuh -- geometry, flat-parameter layout and workspace carve-up of one CRNN configuration.
#pragma once
#include "common.cuh"

namespace sedb200 {

struct Plan {
    int B = 0;
    // conv stack
    int n_conv = 0, C = 0, H = 0;
    int cin[SEDB200_MAX_CONV], win[SEDB200_MAX_CONV], wout[SEDB200_MAX_CONV], pool[SEDB200_MAX_CONV];
    // sequence part
    int T = 0, flat = 0;
    int n_gru = 0, gin[SEDB200_MAX_GRU], gh[SEDB200_MAX_GRU];
    int n_dense = 0, din[SEDB200_MAX_DENSE], dout[SEDB200_MAX_DENSE];
    // flat parameter offsets (floats)
    long conv_w[SEDB200_MAX_CONV], conv_b[SEDB200_MAX_CONV], bn_w[SEDB200_MAX_CONV], bn_b[SEDB200_MAX_CONV];
    long wih[SEDB200_MAX_GRU], whh[SEDB200_MAX_GRU], bih[SEDB200_MAX_GRU], bhh[SEDB200_MAX_GRU];
    long dn_w[SEDB200_MAX_DENSE], dn_b[SEDB200_MAX_DENSE];
    long n_params = 0;
    int n_tensors = 0;
    // workspace offsets (bytes); valid when B > 0
    size_t y[SEDB200_MAX_CONV], stat[SEDB200_MAX_CONV], act[SEDB200_MAX_CONV];   // act[i] = output of block i
    size_t gi[SEDB200_MAX_GRU], gout[SEDB200_MAX_GRU], gates[SEDB200_MAX_GRU];
    size_t gxp[SEDB200_MAX_GRU];              // bf16 hi/lo planes of each GRU layer's input (tensor-core path)
    bool gru_tc[SEDB200_MAX_GRU];
    // scans that write their tensor-core operands themselves (H = 32): h_{t-1} planes per layer (forward scan,
    // kept until the backward pass) and the dgi / dgh planes of the layer being back-propagated
    bool gru_planes[SEDB200_MAX_GRU];
    size_t hpp[SEDB200_MAX_GRU], dgp;
    size_t hp_plane_bytes[SEDB200_MAX_GRU], dg_plane_bytes;
    size_t hid[SEDB200_MAX_DENSE];
    size_t dy, dact[2], dseq[2], dgi, dgh, dhid[2], part, bnsum, tc;
    size_t dys = 0;                           // fp16 gradient-plane scale {s, 1/s} (+ the max |dz| partials behind it)
    // lean block 0 (crnn.cu, "first conv block without its output tensor"): per-window winner bytes written by the
    // fused forward kernel, and the K x (K+1) patch Gram matrix (doubles) both BatchNorm passes are derived from
    size_t arg0 = 0, gram = 0;
    // GRU backward on two streams: weight / bias gradient GEMMs of layer l run on a side stream while the main stream
    // goes on with d(input) and the next scan -- gradient planes and bias partials alternate by layer parity, and the
    // side stream has its own GEMM scratch
    size_t dgp2 = 0, gbias[2] = {0, 0}, tc_side = 0;
    // operand planes of the WEIGHTS, built once per step on the side stream while block 0 runs: conv weights of the
    // plane-native blocks (forward layout and the flipped data-gradient layout) and W_ih of the tensor-core GRU layers
    size_t wpl[SEDB200_MAX_CONV][2], wihp[SEDB200_MAX_GRU];
    bool weight_planes_ahead = false;
    // plane-native tensor-core flow: block i (>= 1) runs fwd, dgrad and wgrad on tcgen05 and exchanges
    // bf16 hi/lo planes with its neighbours instead of fp32 tensors
    bool conv_tc_all[SEDB200_MAX_CONV];
    size_t actp[SEDB200_MAX_CONV];            // planes of block i's OUTPUT (2 x plane, when block i+1 is plane-native)
    size_t dyp;                               // planes of dy (largest plane-native block)
    size_t act_plane_bytes[SEDB200_MAX_CONV], dy_plane_bytes;
    size_t tc_bytes = 0;
    size_t part_floats = 0;
    size_t ws_bytes = 0;
};

// Fused dense head (head_fused.cu): d1 -> relu -> d2 -> sigmoid -> loss and the whole backward of it in one kernel.
bool   head_fused_supported(const Plan& P);
size_t head_fused_part_floats(const Plan& P, int batch);
int    head_fused_run(const Plan& P, const sedb200_crnn_desc* d, const float* params, int batch, const float* x,
                      const float* targets, int loss_kind, float alpha, float gamma, float grad_scale, float* logits,
                      float* probs, float* loss, float* dx, float* grads, float* part, cudaStream_t st);

// Fills `p` from `d` (and the workspace part when batch > 0).  Returns SEDB200_OK or an error code.
int make_plan(const sedb200_crnn_desc* d, int batch, Plan* p);

}  // namespace sedb200
