// crnn_plan.cu -- host-side geometry / layout helpers of the CRNN C ABI (no GPU needed).
#include "crnn_plan.cuh"
#include "gru_scan.cuh"
#include "conv_small.cuh"
#include "gemm_simt.cuh"
#include "tc_conv.cuh"
#include "tc_gemm.cuh"

#include <algorithm>

namespace sedb200 {

namespace {
inline long align_up(long v, long a) { return (v + a - 1) / a * a; }
inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
}  // namespace

int make_plan(const sedb200_crnn_desc* d, int batch, Plan* p) {
    SED_REQUIRE(d != nullptr, SEDB200_EINVAL, "crnn: null descriptor");
    SED_REQUIRE(d->mode == 0 || d->mode == 1, SEDB200_EINVAL, "crnn: mode %d", d->mode);
    SED_REQUIRE(d->in_ch >= 1 && d->H >= 1 && d->W >= 1, SEDB200_EINVAL, "crnn: bad input dims");
    SED_REQUIRE(d->n_conv >= 1 && d->n_conv <= SEDB200_MAX_CONV, SEDB200_EINVAL, "crnn: n_conv %d", d->n_conv);
    SED_REQUIRE(d->n_gru >= 1 && d->n_gru <= SEDB200_MAX_GRU, SEDB200_EINVAL, "crnn: n_gru %d", d->n_gru);
    SED_REQUIRE(d->n_dense >= 1 && d->n_dense <= SEDB200_MAX_DENSE, SEDB200_EINVAL, "crnn: n_dense %d", d->n_dense);
    SED_REQUIRE(d->conv_ch % 4 == 0 && pow2(d->conv_ch / 4) && d->conv_ch <= 1024, SEDB200_ESHAPE,
                "crnn: conv_ch %d must be 4 * a power of two (<= 1024)", d->conv_ch);
    SED_REQUIRE(d->dropout >= 0.0f && d->dropout < 1.0f, SEDB200_EINVAL, "crnn: dropout %f", d->dropout);
    SED_REQUIRE(d->bn_eps > 0.0f, SEDB200_EINVAL, "crnn: bn_eps");
    Plan& P = *p;
    P = Plan();
    P.B = batch;
    P.n_conv = d->n_conv;
    P.C = d->conv_ch;
    P.H = d->H;
    int w = d->W, c = d->in_ch;
    for (int i = 0; i < P.n_conv; ++i) {
        SED_REQUIRE(d->pool[i] >= 1 && d->pool[i] <= 16, SEDB200_EINVAL, "crnn: pool[%d]=%d", i, d->pool[i]);
        P.cin[i] = c;
        P.win[i] = w;
        P.pool[i] = d->pool[i];
        P.wout[i] = w / d->pool[i];
        SED_REQUIRE(P.wout[i] >= 1, SEDB200_ESHAPE, "crnn: pooling leaves no columns after block %d", i);
        w = P.wout[i];
        c = P.C;
    }
    if (d->mode == 0) { P.T = w; P.flat = P.C * P.H; }
    else              { P.T = P.H; P.flat = P.C * w; }
    P.n_gru = d->n_gru;
    int in = P.flat;
    for (int l = 0; l < P.n_gru; ++l) {
        const int h = d->gru_units[l];
        SED_REQUIRE(h >= 4 && h % 4 == 0 && h <= 128, SEDB200_ESHAPE, "crnn: gru_units[%d]=%d (need 4 | h <= 128)", l, h);
        P.gin[l] = in;
        P.gh[l] = h;
        in = 2 * h;
    }
    P.n_dense = d->n_dense;
    for (int j = 0; j < P.n_dense; ++j) {
        SED_REQUIRE(d->dense_units[j] >= 1 && d->dense_units[j] <= 1024, SEDB200_EINVAL, "crnn: dense_units[%d]", j);
        P.din[j] = in;
        P.dout[j] = d->dense_units[j];
        in = P.dout[j];
    }
    // ---- flat parameter layout, every tensor 64-float (256 B) aligned
    long off = 0;
    int nt = 0;
    auto put = [&](long n) { long o = off; off = align_up(off + n, 64); ++nt; return o; };
    for (int i = 0; i < P.n_conv; ++i) {
        P.conv_w[i] = put((long)P.C * P.cin[i] * 9);
        P.conv_b[i] = put(P.C);
        P.bn_w[i] = put(P.C);
        P.bn_b[i] = put(P.C);
    }
    for (int l = 0; l < P.n_gru; ++l) {
        const long h3 = 3L * P.gh[l];
        P.wih[l] = put(2 * h3 * P.gin[l]);
        P.whh[l] = put(2 * h3 * P.gh[l]);
        P.bih[l] = put(2 * h3);
        P.bhh[l] = put(2 * h3);
    }
    for (int j = 0; j < P.n_dense; ++j) {
        P.dn_w[j] = put((long)P.dout[j] * P.din[j]);
        P.dn_b[j] = put(P.dout[j]);
    }
    P.n_params = off;
    P.n_tensors = nt;
    if (batch <= 0) return SEDB200_OK;

    // ---- workspace
    const long B = batch;
    SED_REQUIRE(B * P.H * (long)P.win[0] < (1L << 31) / 16, SEDB200_ESHAPE,
                "crnn: batch %d makes the pixel count of the first block exceed 2^27", batch);
    size_t o = 0;
    auto take = [&](long floats) { size_t at = o; o += (size_t)align_up(floats * 4, 256); return at; };
    long max_y = 0, max_act = 0;
    for (int i = 0; i < P.n_conv; ++i) {
        const long ny = B * P.H * P.win[i] * P.C;
        const long na = B * P.H * P.wout[i] * P.C;
        P.y[i] = take(ny);
        P.stat[i] = take(4L * P.C);
        P.act[i] = take(na);                 // last block: [B][T][flat] (same element count)
        max_y = std::max(max_y, ny);
        max_act = std::max(max_act, na);
    }
    const long BT = B * P.T;
    long max_seq = BT * P.flat, max_g6 = 0;
    for (int l = 0; l < P.n_gru; ++l) {
        P.gi[l] = take(BT * 6 * P.gh[l]);
        P.gout[l] = take(BT * 2 * P.gh[l]);
        P.gates[l] = take(BT * 8 * P.gh[l]);
        max_seq = std::max(max_seq, BT * 2L * P.gh[l]);
        max_g6 = std::max(max_g6, BT * 6L * P.gh[l]);
    }
    long max_hid = 1;
    for (int j = 0; j < P.n_dense; ++j) {
        P.hid[j] = take(BT * P.dout[j]);     // last one unused (logits go to the caller)
        max_hid = std::max(max_hid, BT * (long)P.dout[j]);
    }
    P.dy = take(max_y);
    P.dact[0] = take(max_act);
    P.dact[1] = take(max_act);
    P.dseq[0] = take(max_seq);
    P.dseq[1] = take(max_seq);
    P.dgi = take(max_g6);
    P.dgh = take(max_g6);
    P.dhid[0] = take(max_hid);
    P.dhid[1] = take(max_hid);
    P.bnsum = take(4L * P.C);
    P.dys = take(8 + 1024 + 8 + 2 * SEDB200_MAX_CONV + 8);                  // {scale, 1/scale} + per-block max |dz| partials;
                                                                           // from float 1040: per block {2^b, 2^-(12+b)} (weight fp8 scale)
    P.arg0 = take((B * P.H * P.wout[0] * P.C + 3) / 4);                 // one byte per (window, channel)
    P.gram = take(2L * (9 * P.cin[0]) * (9 * P.cin[0] + 1));           // doubles
    // partial-sum scratch: column-sum partials and split-K partials
    long part = colsum_scratch_floats(B * P.H * P.win[0], std::max({P.C, 6 * 128, 1024}));
    for (int i = 0; i < P.n_conv; ++i)
        part = std::max(part, 64L * P.C * P.cin[i] * 9);                // wgrad: <= 64 K-slices
    for (int l = 0; l < P.n_gru; ++l) {
        part = std::max(part, 64L * 6 * P.gh[l] * std::max(P.gin[l], P.gh[l]));
        part = std::max(part, B * (6L * P.gh[l] * P.gh[l] + 12L * P.gh[l]));         // per-row dW_hh / bias partials
    }
    for (int j = 0; j < P.n_dense; ++j) {
        part = std::max(part, 64L * P.dout[j] * P.din[j]);
        part = std::max(part, ((BT + 127) / 128) * ((long)P.dout[j] * P.din[j] + P.dout[j]));     // per-block dW/db partials
    }
    part = std::max(part, B * ((P.H + 7) / 8) * (long)std::max(2, P.cin[0] * 9 + 1) * P.C     // direct block-0 partials
                              + 2L * (P.cin[0] * 9 + 1) * P.C + 64);                               // + their column sums (doubles)
    part = std::max(part, 148L * 16 * 2 * P.C);
    for (int i = 1; i < P.n_conv; ++i) part = std::max(part, 2L * P.C * B * ((P.H * P.win[i] + 127) / 128 + 1));   // conv-epilogue BN partials                                                 // BN backward sums
    if (head_fused_supported(P)) part = std::max(part, (long)head_fused_part_floats(P, (int)B));   // fused head partials
    for (int i = 0; i < P.n_conv; ++i)                                                            // small-channel wgrad partials
        if (conv_small_wgrad_supported(P.cin[i], P.C, P.win[i]))
            part = std::max(part, (long)conv_small_wgrad_part_floats(P.cin[i], P.C, (int)B, P.H));
    P.part_floats = (size_t)part;
    P.part = take(part);
    // plane-native tensor-core blocks
    auto plane = [](long n) { return ((size_t)n * 2 + 1023) & ~(size_t)1023; };
    size_t dyp_bytes = 0;
    for (int i = 0; i < P.n_conv; ++i) { P.conv_tc_all[i] = false; P.actp[i] = 0; P.act_plane_bytes[i] = 0; }
    for (int i = 1; i < P.n_conv; ++i)
        P.conv_tc_all[i] = d->tensor_cores && conv_tc_supported(P.H, P.win[i], P.cin[i], P.C) &&
                           conv_tc_supported(P.H, P.win[i], P.C, P.cin[i]) && wgrad_tc_supported(P.H, P.win[i], P.cin[i], P.C);
    for (int i = 0; i + 1 < P.n_conv; ++i)
        if (P.conv_tc_all[i + 1]) {
            P.act_plane_bytes[i] = plane(B * P.H * P.wout[i] * P.C);
            P.actp[i] = take((long)(2 * P.act_plane_bytes[i] / 4));
        }
    for (int i = 1; i < P.n_conv; ++i)
        if (P.conv_tc_all[i]) dyp_bytes = std::max(dyp_bytes, plane(B * P.H * P.win[i] * P.C));
    P.dy_plane_bytes = dyp_bytes;
    P.dyp = take((long)(2 * dyp_bytes / 4) + 64);
    // tensor-core scratch: bf16 operand planes + wgrad split-K partials
    size_t tc = 0;
    for (int i = 1; i < P.n_conv; ++i)
        if (P.conv_tc_all[i])
            tc = std::max(tc, conv_tc_weight_scratch_bytes(P.cin[i], P.C) + wgrad_tc_part_bytes(P.cin[i], P.C) + 4096);
    if (d->tensor_cores)
        for (int i = 1; i < P.n_conv; ++i) {
            if (conv_tc_supported(P.H, P.win[i], P.cin[i], P.C)) tc = std::max(tc, conv_tc_scratch_bytes(batch, P.H, P.win[i], P.cin[i], P.C));
            if (conv_tc_supported(P.H, P.win[i], P.C, P.cin[i])) tc = std::max(tc, conv_tc_scratch_bytes(batch, P.H, P.win[i], P.C, P.cin[i]));
            if (wgrad_tc_supported(P.H, P.win[i], P.cin[i], P.C)) tc = std::max(tc, wgrad_tc_scratch_bytes(batch, P.H, P.win[i], P.cin[i], P.C));
        }
    for (int l = 0; l < P.n_gru; ++l) {
        const int h6 = 6 * P.gh[l], in = P.gin[l];
        P.gru_tc[l] = d->tensor_cores && BT >= 1024 && gemm_tc_supported((int)BT, h6, in) && gemm_tc_supported(h6, in, (int)BT) &&
                      gemm_tc_supported((int)BT, in, h6);
        P.gxp[l] = 0;
        if (!P.gru_tc[l]) continue;
        P.gxp[l] = take((long)(2 * plane(BT * in) / 4));
        tc = std::max(tc, 2 * plane(BT * h6) + 2 * plane((long)h6 * in) + (size_t)sm_count() * h6 * in * 4 + 4096);
        tc = std::max(tc, 2 * plane(BT * h6) + 2 * plane(BT * 2L * P.gh[l]) + ((size_t)sm_count() + 1) * h6 * 2 * P.gh[l] * 4 + 8192);
    }
    P.dg_plane_bytes = 0;
    for (int l = 0; l < P.n_gru; ++l) {
        P.gru_planes[l] = P.gru_tc[l] && gru_scan_emits_planes(P.gh[l]) && gemm_tc_supported(6 * P.gh[l], 2 * P.gh[l], (int)BT);
        P.hpp[l] = 0;
        P.hp_plane_bytes[l] = 0;
        if (!P.gru_planes[l]) continue;
        P.hp_plane_bytes[l] = plane(BT * 2L * P.gh[l]);
        P.hpp[l] = take((long)(2 * P.hp_plane_bytes[l] / 4));
        P.dg_plane_bytes = std::max(P.dg_plane_bytes, plane(BT * 6L * P.gh[l]));
    }
    P.dgp = P.dg_plane_bytes ? take((long)(4 * P.dg_plane_bytes / 4)) : 0;
    if (P.dg_plane_bytes) {
        P.dgp2 = take((long)(4 * P.dg_plane_bytes / 4));
        size_t side = 0;
        long gb = 0;
        for (int l = 0; l < P.n_gru; ++l) {
            if (!P.gru_planes[l]) continue;
            const size_t h6 = 6 * (size_t)P.gh[l];
            side = std::max(side, ((size_t)sm_count() + 1) * h6 * 2 * P.gh[l] * 4 + 8192);      // dW_hh: tmp + split-K partials
            side = std::max(side, (size_t)sm_count() * h6 * P.gin[l] * 4 + 4096);                // dW_ih split-K partials
            gb = std::max(gb, B * 24L * P.gh[l]);
        }
        P.gbias[0] = take(gb);
        P.gbias[1] = take(gb);
        P.tc_side = take((long)(side / 4) + 64);
    }
    for (int i = 0; i < P.n_conv; ++i) {
        P.wpl[i][0] = P.wpl[i][1] = 0;
        if (!P.conv_tc_all[i]) continue;
        P.wpl[i][0] = take((long)(conv_tc_weight_scratch_bytes(P.cin[i], P.C) / 4));
        P.wpl[i][1] = take((long)(conv_tc_weight_scratch_bytes(P.cin[i], P.C) / 4));
        P.weight_planes_ahead = true;
    }
    for (int l = 0; l < P.n_gru; ++l) {
        P.wihp[l] = 0;
        if (!P.gru_tc[l]) continue;
        P.wihp[l] = take((long)(2 * plane(6L * P.gh[l] * P.gin[l]) / 4));
        P.weight_planes_ahead = true;
    }
    P.tc_bytes = tc;
    P.tc = take((long)(tc / 4) + 64);
    P.ws_bytes = o;
    return SEDB200_OK;
}

}  // namespace sedb200

using namespace sedb200;

extern "C" {

int sedb200_crnn_validate(const sedb200_crnn_desc* d) {
    Plan p;
    return make_plan(d, 0, &p);
}
int sedb200_crnn_seq_len(const sedb200_crnn_desc* d) {
    Plan p;
    return make_plan(d, 0, &p) ? -1 : p.T;
}
int sedb200_crnn_flat(const sedb200_crnn_desc* d) {
    Plan p;
    return make_plan(d, 0, &p) ? -1 : p.flat;
}
int sedb200_crnn_n_tensors(const sedb200_crnn_desc* d) {
    Plan p;
    return make_plan(d, 0, &p) ? -1 : p.n_tensors;
}
long sedb200_crnn_param_layout(const sedb200_crnn_desc* d, long* offsets) {
    Plan p;
    if (make_plan(d, 0, &p)) return -1;
    if (offsets) {
        int k = 0;
        for (int i = 0; i < p.n_conv; ++i) {
            offsets[k++] = p.conv_w[i];
            offsets[k++] = p.conv_b[i];
            offsets[k++] = p.bn_w[i];
            offsets[k++] = p.bn_b[i];
        }
        for (int l = 0; l < p.n_gru; ++l) {
            offsets[k++] = p.wih[l];
            offsets[k++] = p.whh[l];
            offsets[k++] = p.bih[l];
            offsets[k++] = p.bhh[l];
        }
        for (int j = 0; j < p.n_dense; ++j) {
            offsets[k++] = p.dn_w[j];
            offsets[k++] = p.dn_b[j];
        }
    }
    return p.n_params;
}
long sedb200_crnn_bn_state_floats(const sedb200_crnn_desc* d) {
    Plan p;
    return make_plan(d, 0, &p) ? -1 : 2L * p.n_conv * p.C;
}
size_t sedb200_crnn_workspace_bytes(const sedb200_crnn_desc* d, int batch) {
    Plan p;
    if (batch <= 0 || make_plan(d, batch, &p)) return 0;
    return p.ws_bytes;
}

}  // extern "C"
