"""One tiny CRNN training step on cuda:0, checked against the CPU oracle (called by __graft_entry__.smoke)."""
from __future__ import annotations

from dataclasses import replace


def run() -> None:
    import torch
    from oracle import crnn_ref as R
    from . import config, engine

    # (a) the tree's own configuration (fp32 CUDA-core path), (b) a 128-channel SEDnet slice (tcgen05 path)
    for preset, ov, batch, loss in (("fork", {}, 16, "focal"), ("c2", {"seq_len": 32}, 4, "bce")):
        rcfg = {**R.PRESETS[preset], **ov}
        torch.manual_seed(0)
        ref = R.RefCRNN(**rcfg)
        eng = engine.CRNNEngine(replace(config.PRESETS[preset], dropout=0.0, **ov), loss=loss, weight_decay=1e-4, clip=1.0)
        eng.load_named({k: v.detach() for k, v in ref.canonical_named_params()})
        x, y = R.synth_batch(rcfg, batch, seed=1)
        opt = R.make_adam(ref, 1e-3, 1e-4)
        loss_ref, logits_ref, _ = R.train_step(ref, opt, x, y, loss, 1.0)
        with torch.no_grad():
            p_ref = torch.sigmoid(ref(x))
        l, _ = eng.train_step(x.cuda(), y.cuda())
        p = eng.predict_proba(x.cuda(), training_bn=True).cpu()
        err = (p - p_ref).abs().max().item()
        assert abs(l.item() - loss_ref.item()) < 1e-5, (l.item(), loss_ref.item())
        assert err <= 1e-3, err
        safe = (p_ref - 0.5).abs() > 2 * err + 1e-7
        assert torch.equal((p > 0.5)[safe], (p_ref > 0.5)[safe])
        print(f"smoke: CRNN[{preset}{ov or ''}] one train step: loss {l.item():.6f} (oracle {loss_ref.item():.6f}), "
              f"max |dp| after the step {err:.2e} (gate 1e-3), decisions identical")
