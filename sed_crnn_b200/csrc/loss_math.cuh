// loss_math.cuh -- per-element loss and its derivative w.r.t. the logit, shared by the stand-alone loss kernel
// (head_optim.cu) and the fused dense head (head_fused.cu).
//   FocalBCELoss        crnn_lightning.py:27-35 (EPS crnn_lightning.py:21)
//   BCEWithLogitsLoss   sed.py:160
#pragma once
#include "common.cuh"

namespace sedb200 {

// p = sigmoid(l); loss = element loss; dl = d loss / d l  (unscaled: the caller applies 1/n and the grad scale)
__device__ __forceinline__ void loss_elem(int kind, float alpha, float gamma, float l, float t, float& p, float& loss,
                                          float& dl) {
    p = 1.0f / (1.0f + expf(-l));
    if (kind == SEDB200_LOSS_FOCAL) {
        const bool pos = (t == 1.0f);
        const float pt = pos ? p : 1.0f - p;
        const float om = 1.0f - pt;
        const float lg = logf(pt + 1e-12f);
        const float pw = powf(om, gamma);
        loss = -alpha * pw * lg;
        // d loss / d pt = alpha * gamma * om^(gamma-1) * lg - alpha * om^gamma / (pt + eps)
        const float pw1 = (gamma == 2.0f) ? om : powf(om, gamma - 1.0f);
        const float dpt = alpha * gamma * pw1 * lg - alpha * pw / (pt + 1e-12f);
        dl = dpt * (pos ? 1.0f : -1.0f) * p * (1.0f - p);
    } else {
        loss = fmaxf(l, 0.0f) - l * t + log1pf(expf(-fabsf(l)));
        dl = p - t;
    }
}

}  // namespace sedb200
