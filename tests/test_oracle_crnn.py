"""Pin oracle/crnn_ref.py to the reference's own modules via the golden fixtures written by
oracle/make_golden.py (which executed crnn_lightning.py / sed.py unmodified on torch-CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import crnn_ref as R

LIGHTNING_MAP = {  # reference state_dict key -> oracle canonical name
    **{f"conv_stack.{4 * i}.{p}": f"conv{i}.{p}" for i in range(3) for p in ("weight", "bias")},
    **{f"conv_stack.{4 * i + 1}.{p}": f"bn{i}.{p}" for i in range(3) for p in ("weight", "bias")},
    **{f"gru{i + 1}.{a}_l0{sfx}": f"gru{i}.{tag}.{b}" for i in range(2)
       for sfx, tag in (("", "f"), ("_reverse", "r"))
       for a, b in (("weight_ih", "w_ih"), ("weight_hh", "w_hh"), ("bias_ih", "b_ih"), ("bias_hh", "b_hh"))},
    "d1.weight": "dense0.weight", "d1.bias": "dense0.bias", "d2.weight": "dense1.weight", "d2.bias": "dense1.bias",
}
SEDPY_MAP = {
    **{f"convs.{i}.{p}": f"conv{i}.{p}" for i in range(3) for p in ("weight", "bias")},
    **{f"bns.{i}.{p}": f"bn{i}.{p}" for i in range(3) for p in ("weight", "bias")},
    **{f"gru.{a}_l{i}{sfx}": f"gru{i}.{tag}.{b}" for i in range(2)
       for sfx, tag in (("", "f"), ("_reverse", "r"))
       for a, b in (("weight_ih", "w_ih"), ("weight_hh", "w_hh"), ("bias_ih", "b_ih"), ("bias_hh", "b_hh"))},
    "fc.weight": "dense0.weight", "fc.bias": "dense0.bias",
}


def load_into(model, g, keymap, prefix="w."):
    named = dict(model.canonical_named_params())
    with torch.no_grad():
        for rk, ck in keymap.items():
            named[ck].copy_(torch.from_numpy(g[prefix + rk]))


@pytest.mark.parametrize("fixture,preset,keymap,loss,wd,clip", [
    ("crnn_fork_lightning.npz", "fork", LIGHTNING_MAP, "focal", 1e-4, 1.0),
    ("crnn_fork_sedpy.npz", "sedpy", SEDPY_MAP, "bce", 0.0, None),
])
def test_oracle_equals_reference_one_step(golden_dir, fixture, preset, keymap, loss, wd, clip):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, fixture))
    cfg = dict(R.PRESETS[preset])
    if preset == "sedpy":
        cfg["conv_ch"] = 32
    m = R.RefCRNN(**cfg)
    load_into(m, g, keymap)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    opt = R.make_adam(m, 1e-3, wd)
    loss0, logits0, gnorm = R.train_step(m, opt, x, y, loss, clip)
    np.testing.assert_allclose(logits0.numpy(), g["logits0"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(loss0.item(), g["loss0"], rtol=1e-6)
    if clip:
        np.testing.assert_allclose(gnorm.item(), g["gnorm"], rtol=1e-5)
    named = dict(m.canonical_named_params())
    for rk, ck in keymap.items():
        if preset == "sedpy" and ck.startswith("conv") and ck.endswith("bias"):
            continue            # true gradient 0 (BN follows); Adam amplifies rounding noise
        np.testing.assert_allclose(named[ck].detach().numpy(), g["w1." + rk], rtol=0, atol=2e-6, err_msg=rk)
    with torch.no_grad():
        p1 = torch.sigmoid(m(x)).numpy()
    np.testing.assert_allclose(p1, g["probs1"], rtol=0, atol=1e-6)


def test_loss_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_kat.npz"))
    lo, t = torch.from_numpy(g["logits"]), torch.from_numpy(g["targets"])
    assert abs(R.focal_bce(lo, t).item() - float(g["focal_mean"])) < 1e-7
    assert abs(R.focal_bce(lo, t, reduction="sum").item() - float(g["focal_sum"])) < 1e-5
    assert abs(R.focal_bce(lo, t, 0.5, 1.0).item() - float(g["focal_a5_g1"])) < 1e-7
    assert abs(R.bce_logits(lo, t).item() - float(g["bce_mean"])) < 1e-7
    assert abs(float(g["focal_mean"]) - 0.1736757904) < 1e-9 and abs(float(g["bce_mean"]) - 1.0583852530) < 1e-9


def test_sednet_mode_shapes():
    for name in ("c1", "c2"):
        cfg = dict(R.PRESETS[name])
        cfg["seq_len"] = 32
        m = R.RefCRNN(**cfg)
        x, y = R.synth_batch(cfg, 2)
        assert m(x).shape == y.shape == (2, 32, 6)
        assert m.flat == 128 * 2
