#!/usr/bin/env python3
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference files from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Fixtures written (all small, committed):
  metrics_kat.npz        inputs + outputs of reference metrics.py (f1/er framewise + 1sec)
  loss_kat.npz           reference FocalBCELoss / BCEWithLogitsLoss on a fixed logit ramp
  crnn_fork_lightning.npz  crnn_lightning.TimePooledCRNN: weights, batch, logits, loss, grads,
                           and probabilities after one clip(1.0)+Adam(wd 1e-4) step
  crnn_fork_sedpy.npz      sed.TimePooledCRNN(conv_channels=32): same, BCE, Adam(wd 0), no clip
  window_sampler.npz       decorte_datamodule.HitWindowDataset (plain + SpecAugment) items, _find_clean_negatives
  logmel_oracle.npz        NOT from the reference (librosa absent): frozen output of
                           oracle/logmel_ref.py, a regression anchor only (parity unpinned)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, logmel_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def metrics_kat():
    m = ref_import.load("metrics")
    cases = {}
    rng = np.random.default_rng(7)

    def add(name, O, T, block):
        with np.errstate(all="ignore"):
            vals = np.array([m.f1_overall_framewise(O, T), m.er_overall_framewise(O, T),
                             m.f1_overall_1sec(O, T, block), m.er_overall_1sec(O, T, block)],
                            dtype=np.float64)
        cases[name + "_O"], cases[name + "_T"] = O, T
        cases[name + "_block"], cases[name + "_out"] = np.int64(block), vals

    i = np.arange(7 * 8 * 1)
    add("m1", ((7 * i + 3) % 5 < 2).astype(np.uint8).reshape(7, 8, 1),
        ((3 * i + 1) % 4 == 0).astype(np.uint8).reshape(7, 8, 1), 5)           # SURVEY KAT-M1
    i = np.arange(3 * 256 * 6)
    add("m2", ((11 * i + 5) % 13 < 3).astype(np.uint8).reshape(3, 256, 6),
        ((5 * i + 2) % 9 < 2).astype(np.uint8).reshape(3, 256, 6), 43)         # SURVEY KAT-M2
    add("rand_bool", rng.random((5, 64, 6)) > 0.7, (rng.random((5, 64, 6)) > 0.8).astype(np.float32), 50)
    add("rand_big", (rng.random((64, 8, 1)) > 0.5).astype(np.uint8),
        (rng.random((64, 8, 1)) > 0.6).astype(np.float32), 5)                  # fork epoch-end shape
    add("no_ref", (rng.random((4, 8, 1)) > 0.5).astype(np.uint8), np.zeros((4, 8, 1), np.uint8), 5)
    add("all_zero", np.zeros((4, 8, 2), np.uint8), np.zeros((4, 8, 2), np.uint8), 5)
    add("perfect", (rng.random((4, 16, 3)) > 0.5).astype(np.uint8), None, 5) if False else None
    P = (rng.random((4, 16, 3)) > 0.5).astype(np.uint8)
    add("perfect", P, P.copy(), 5)
    add("short", (rng.random((1, 3, 2)) > 0.5).astype(np.uint8), (rng.random((1, 3, 2)) > 0.5).astype(np.uint8), 5)
    add("twod", (rng.random((100, 6)) > 0.6).astype(np.uint8), (rng.random((100, 6)) > 0.6).astype(np.uint8), 7)
    cases["names"] = np.array(["m1", "m2", "rand_bool", "rand_big", "no_ref", "all_zero", "perfect", "short", "twod"])
    np.savez_compressed(os.path.join(OUT, "metrics_kat.npz"), **cases)
    print("metrics_kat: m1", cases["m1_out"], "m2", cases["m2_out"])


def loss_kat():
    cl = ref_import.load("crnn_lightning")
    logits = torch.linspace(-3, 3, 32)
    t = (torch.arange(32) % 3 == 0).float()
    out = dict(logits=logits.numpy(), targets=t.numpy(),
               focal_mean=cl.FocalBCELoss()(logits, t).item(),
               focal_sum=cl.FocalBCELoss(reduction="sum")(logits, t).item(),
               focal_a5_g1=cl.FocalBCELoss(alpha=.5, gamma=1.)(logits, t).item(),
               bce_mean=torch.nn.BCEWithLogitsLoss()(logits, t).item())
    np.savez(os.path.join(OUT, "loss_kat.npz"), **out)
    print("loss_kat:", {k: v for k, v in out.items() if not hasattr(v, "shape")})


def _one_step(model, loss_fn, x, y, *, wd, clip):
    model.train()
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    logits = model(x)
    loss = loss_fn(logits, y)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=wd)
    opt.zero_grad()
    loss.backward()
    grads = {"grad." + k: p.grad.detach().clone().numpy() for k, p in model.named_parameters()}
    gnorm = np.float64(torch.nn.utils.clip_grad_norm_(model.parameters(), clip).item()) if clip else np.float64(
        torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters())).item())
    opt.step()
    sd_mid = {"after_fwd." + k: v.detach().clone().numpy() for k, v in model.state_dict().items()
              if "running" in k}
    with torch.no_grad():
        logits1 = model(x)                       # train-mode BN, dropout 0
    out = {"w." + k: v for k, v in sd0.items()}
    out.update(grads)
    out.update(sd_mid)
    out.update({"w1." + k: v.detach().clone().numpy() for k, v in model.named_parameters()})
    out.update(x=x.numpy(), y=y.numpy(), logits0=logits.detach().numpy(), loss0=np.float64(loss.item()),
               gnorm=gnorm, probs1=torch.sigmoid(logits1).numpy())
    model.eval()
    return out


def crnn_fork():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    cl = ref_import.load("crnn_lightning")
    g = torch.Generator().manual_seed(123)
    x = torch.randn(16, 1, 40, 64, generator=g)
    y = (torch.rand(16, 8, 1, generator=g) < 0.2).float()
    model = cl.TimePooledCRNN(dropout=0.0)
    out = _one_step(model, cl.FocalBCELoss(), x, y, wd=1e-4, clip=1.0)
    with torch.no_grad():
        out["logits_eval"] = model(x).numpy()    # eval-mode BN (running stats after 2 train fwd)
    np.savez_compressed(os.path.join(OUT, "crnn_fork_lightning.npz"), **out)
    print("crnn_fork_lightning: loss0", out["loss0"], "gnorm", out["gnorm"],
          "margin", np.abs(out["probs1"] - 0.5).min())

    torch.manual_seed(1)
    sed = ref_import.load("sed")
    model = sed.TimePooledCRNN(conv_channels=32, dropout=0.0)
    out = _one_step(model, torch.nn.BCEWithLogitsLoss(), x, y, wd=0.0, clip=None)
    np.savez_compressed(os.path.join(OUT, "crnn_fork_sedpy.npz"), **out)
    print("crnn_fork_sedpy: loss0", out["loss0"], "gnorm", out["gnorm"],
          "margin", np.abs(out["probs1"] - 0.5).min())


def logmel_anchor():
    out = {}
    for name, n, ch, kind in (("mix_1s", 44100, 1, "mix"), ("noise_odd", 2 * 44100 + 1, 1, "noise"),
                              ("chirp_stereo", 30000, 2, "chirp"), ("short", 1000, 1, "mix")):
        y = logmel_ref.synth_clip(hash(name) % 1000 if False else len(name), n, ch, kind)
        out[name + "_pcm"] = y
        for pm in ("constant", "reflect"):
            out[f"{name}_{pm}"] = np.ascontiguousarray(logmel_ref.mbe_multichannel(y, pad_mode=pm))
    out["mel_fb"] = logmel_ref.mel_filterbank()
    np.savez_compressed(os.path.join(OUT, "logmel_oracle.npz"), **out)
    print("logmel_oracle:", {k: v.shape for k, v in out.items() if "_pcm" not in k})


def window_sampler():
    """HitWindowDataset of the UNMODIFIED reference decorte_datamodule.py under fixed seeds."""
    import random
    dm = ref_import.load("decorte_datamodule")
    rng = np.random.default_rng(11)
    n = 900
    mel = rng.standard_normal((n, 40)).astype(np.float32)
    lab = np.zeros((n, 1), np.float32)
    for a, b in ((70, 75), (200, 203), (204, 260), (500, 501), (880, 900)):
        lab[a:b, 0] = 1.0
    out = dict(mel=mel, lab=lab, neg_starts=np.asarray(dm._find_clean_negatives(lab), dtype=np.int64))
    idx = np.arange(24)
    for name, aug in (("plain", False), ("aug", True)):
        random.seed(5)
        np.random.seed(6)
        ds = dm.HitWindowDataset(mel, lab, augment=aug)
        items = [ds[int(i)] for i in idx]
        out[f"{name}_x"] = np.stack([it[0].numpy() for it in items])
        out[f"{name}_y"] = np.stack([it[1].numpy() for it in items])
        out[f"{name}_len"] = np.int64(len(ds))
        out[f"{name}_pos_frames"] = np.asarray(ds.pos_frames, dtype=np.int64)
    out["idx"] = idx
    out["seeds"] = np.array([5, 6])
    np.savez_compressed(os.path.join(OUT, "window_sampler.npz"), **out)
    print("window_sampler:", out["plain_x"].shape, out["plain_y"].shape, "neg starts", out["neg_starts"].size,
          "masked zeros", int((out["aug_x"] == 0).sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-window" in sys.argv:
        window_sampler()
        sys.exit(0)
    metrics_kat()
    loss_kat()
    crnn_fork()
    logmel_anchor()
    window_sampler()
