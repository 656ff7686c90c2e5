"""Tiny end-to-end exercise of every kernel family (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dataclasses import replace
import numpy as np, torch
from oracle import logmel_ref, crnn_ref as R
from sed_crnn_b200 import feature, config, engine

y = logmel_ref.synth_clip(0, 5000, 2, "mix")
for pm in ("constant", "reflect"):
    out = feature.mbe_device(torch.from_numpy(y).cuda(), pad_mode=pm)
print("logmel ok", tuple(out.shape))
for preset, ov, b in (("fork", {}, 4), ("c2", {"seq_len": 16}, 2), ("c5", {"seq_len": 16, "gru_units": (128, 16)}, 1)):
    rcfg = {**R.PRESETS[preset], **ov}
    eng = engine.CRNNEngine(replace(config.PRESETS[preset], **ov), loss="bce")
    eng.init_default(0)
    x, t = R.synth_batch(rcfg, b, seed=1)
    for _ in range(2):
        loss, probs = eng.train_step(x.cuda(), t.cuda())
    c = eng.threshold_counts(probs, t.cuda(), 5)
    torch.cuda.synchronize()
    print(preset, "ok", loss.item(), c.tolist()[:3])
