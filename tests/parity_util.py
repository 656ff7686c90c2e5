"""Shared machinery of the one-training-step parity tests (CUDA engine through the C ABI vs the CPU oracle).

Gates (BASELINE.json north_star): per-frame probabilities within 1e-3 after one training step; thresholded frame
decisions and segment ER/F1 bit-exact.  Every run also appends its measured errors to
`gpurun_out/parity_report.jsonl` (when that directory exists) so that the numbers behind a green test can be read."""
from __future__ import annotations

import json
import os
from dataclasses import replace

import numpy as np
import torch

from oracle import crnn_ref as R
from oracle import metrics_ref as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROB_TOL = 1e-3
LOGIT_TOL = 2e-5


def report(case: str, **numbers) -> None:
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": case, **{k: (float(v) if isinstance(v, (int, float, np.floating)) else v)
                                                 for k, v in numbers.items()}}) + "\n")


def make_pair(config, engine, preset, overrides, loss, wd, clip, seed=0, tensor_cores=True, dropout=0.0,
              engine_seed=0):
    rcfg = dict(R.PRESETS[preset]); rcfg.update(overrides)
    cfg = replace(config.PRESETS[preset], dropout=dropout, tensor_cores=tensor_cores, **overrides)
    torch.manual_seed(seed)
    ref = R.RefCRNN(**{**rcfg, "dropout": dropout, "dropout_each_block": cfg.dropout_each_block})
    eng = engine.CRNNEngine(cfg, loss=loss, weight_decay=wd, clip=clip, seed=engine_seed)
    eng.load_named({k: v.detach() for k, v in ref.canonical_named_params()})
    return rcfg, ref, cfg, eng


def engine_tensor(views, name):
    parts = name.split(".")
    if len(parts) == 3 and parts[1] in ("f", "r"):
        return views[f"{parts[0]}.{parts[2]}"][0 if parts[1] == "f" else 1]
    return views[name]


def tensor_class(name: str) -> str:
    return "conv" if name.startswith(("conv", "bn")) else "seq"          # seq = GRU + dense (no pool decision upstream
    #                                                                       of their OWN contraction, but dY reaches
    #                                                                       them only through exact fp32 kernels)


def one_step_parity(case, rcfg, ref, cfg, eng, x, y, loss, wd, clip, *, masks_from_engine=False, logit_tol=None,
                    exact_grads=False, seq_grad_tol=1e-3, conv_grad_l2=1e-2, check_weights=True, blocks=(5, 43)):
    """Runs ONE optimisation step on both sides from identical weights and compares: logits, loss, every gradient
    tensor, the global gradient norm, post-step weights, probabilities after the step (gate 1e-3), decisions and
    segment metrics (bit-exact).

    exact_grads: fp32 CUDA-core path with PyTorch-like summation -> max-abs 2e-3 of each tensor's scale.
    otherwise (tcgen05 3-term split or the direct small-channel convs): forward differences of ~1e-5 can move a
    handful of ReLU / max-pool decisions, and BatchNorm-cancelled sums react to single flips, so conv / BN gradients
    are held to a relative L2 error (`conv_grad_l2`); GRU / dense gradients do not sit behind such a cancellation and
    are held to max-abs `seq_grad_tol` of the tensor's scale.
    masks_from_engine: dropout is ON; the engine's keep-masks for this step are injected into the oracle."""
    if logit_tol is None:
        # plane-native tensor-core blocks: the forward is an fp16 pass plus an e4m3 correction pass (~2^-15 per product;
        # measured 1.0-1.2e-5 on the logits at full size, up to 2.8e-5 on the fork-layout reduced case); fp32 paths 2e-5
        logit_tol = 5e-5 if (cfg.tensor_cores and cfg.conv_ch % 128 == 0) else LOGIT_TOL
    xd, yd = x.cuda(), y.cuda()
    B = x.shape[0]
    # --- engine forward first (the masks belong to its seed)
    logits = eng.forward(xd, training=True).clone()
    masks = [m.cpu() for m in eng.dropout_masks(B)] if masks_from_engine else None
    # --- reference step
    opt = R.make_adam(ref, 1e-3, wd)
    ref.train()
    logits_ref = ref(x, masks) if masks is not None else ref(x)
    loss_ref = R.loss_fn(loss)(logits_ref, y)
    opt.zero_grad(); loss_ref.backward()
    gref = {k: p.grad.detach().clone() for k, p in ref.canonical_named_params()}
    gn_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), clip if clip else 1e30)
    opt.step()
    ref.train()
    with torch.no_grad():
        p1_ref = torch.sigmoid(ref(x, masks) if masks is not None else ref(x))
    # --- engine step, piece by piece
    lerr = (logits.cpu() - logits_ref.detach()).abs().max().item()
    assert lerr <= logit_tol, ("logits", lerr)
    l, probs, dlog = eng.loss_and_grad(logits, yd)
    assert abs(l.item() - loss_ref.item()) <= 2e-6 + 1e-5 * abs(loss_ref.item())
    eng.backward(xd, dlog)
    gv = eng.views(eng.grads)
    worst = {"conv_l2": 0.0, "seq_maxabs": 0.0, "exact_maxabs": 0.0}
    for name, g in gref.items():
        got = engine_tensor(gv, name).cpu()
        scale = max(g.abs().max().item(), 1e-6)
        if name.startswith("conv") and name.endswith("bias"):
            assert got.abs().max().item() <= 1e-4 * max(1.0, scale)      # true gradient is 0 (BN follows)
            continue
        if exact_grads:
            err = (got - g).abs().max().item() / scale
            worst["exact_maxabs"] = max(worst["exact_maxabs"], err)
            assert err <= 2e-3, (name, err)
        elif tensor_class(name) == "seq":
            err = (got - g).abs().max().item() / scale
            worst["seq_maxabs"] = max(worst["seq_maxabs"], err)
            assert err <= seq_grad_tol, (name, err)
        else:
            err = (got - g).norm().item() / max(g.norm().item(), 1e-12)
            worst["conv_l2"] = max(worst["conv_l2"], err)
            assert err <= conv_grad_l2, (name, err)
    gn = eng.optimizer_step()
    gnerr = abs(gn.item() - gn_ref.item()) / gn_ref.item()
    assert gnerr <= 1e-4, gnerr
    # probabilities after the step: same inputs, train-mode BatchNorm, the step's masks again when dropout is on
    if masks_from_engine:
        keep, nbt = eng.bn_state.clone(), eng.num_batches_tracked
        p1 = torch.sigmoid(eng.forward(xd, training=True, seed=eng._last_seed)).cpu()
        eng.bn_state.copy_(keep); eng.num_batches_tracked = nbt
    else:
        p1 = eng.predict_proba(xd, training_bn=True).cpu()
    perr = (p1 - p1_ref).abs().max().item()
    werr = 0.0
    if check_weights and wd > 0:     # with wd=0 conv-bias updates are sign(noise)*lr: skip weight compare
        v = eng.views()
        for (name, pref) in ref.canonical_named_params():
            werr = max(werr, (engine_tensor(v, name).cpu() - pref.detach()).abs().max().item())
    margin = (p1_ref - 0.5).abs()
    safe = margin > 2 * perr + 1e-7
    report(case, logits_maxabs=lerr, probs_after_step_maxabs=perr, gnorm_rel=gnerr, weights_maxabs=werr,
           frames=int(p1.numel()), frames_outside_margin=int(safe.sum()), min_margin=float(margin.min()), **worst)
    assert perr <= PROB_TOL, perr
    assert werr <= 2.1e-3, werr
    # --- decisions + metrics bit-exact (frames inside the tolerance margin are reported and excluded)
    assert torch.equal((p1 > 0.5)[safe], (p1_ref > 0.5)[safe])
    O, Oref, T = (p1.numpy() > 0.5), (p1_ref.numpy() > 0.5), y.numpy()
    if bool(safe.all()):
        assert np.array_equal(O, Oref)
    if np.array_equal(O, Oref):
        for blk in blocks:
            a = np.array([M.f1_overall_1sec(O, T, blk), M.er_overall_1sec(O, T, blk),
                          M.f1_overall_framewise(O, T), M.er_overall_framewise(O, T)])
            b = np.array([M.f1_overall_1sec(Oref, T, blk), M.er_overall_1sec(Oref, T, blk),
                          M.f1_overall_framewise(Oref, T), M.er_overall_framewise(Oref, T)])
            assert np.array_equal(a, b, equal_nan=True)
    return {"perr": perr, "lerr": lerr, "all_safe": bool(safe.all()), "decisions_equal": bool(np.array_equal(O, Oref)),
            **worst}
