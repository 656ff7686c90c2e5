"""Parity of the BASELINE.json configurations AT THEIR OWN SIZE, tensor cores on, through the C ABI, against the CPU
oracle (oracle/crnn_ref.py: the stock torch.nn modules the reference instantiates, crnn_lightning.py:41-73 /
sed.py:82-112), plus dropout checked against the oracle by mask injection.

What full size exercises that the reduced cases of test_crnn_gpu.py do not: the 148-CTA persistent tile schedule of
conv_tc_kernel with > 1 tile per CTA inside the whole network, the split-K slice counts of wgrad_tc_kernel at
M = 262,144, BatchNorm partial-sum finalizers over 2,048 M-tiles, GRU recurrences of 256 and 2,048 steps.

The oracle step (forward + backward + Adam + a second forward) takes a few seconds per case on the box's host cores."""
import os

import numpy as np
import pytest
import torch

import parity_util as PU
from oracle import crnn_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built_lib):
    assert torch.cuda.is_available()
    from sed_crnn_b200 import config, engine
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return config, engine


@pytest.mark.parametrize("preset,seed", [("c1", 0), ("c2", 0), ("c2", 1)])
def test_c1_c2_full_size_training_step(pkg, preset, seed):
    """BASELINE configs[0] (mono) and configs[1] (binaural): batch 128, seq_len 256, 128 filters, 6 classes."""
    config, engine = pkg
    rcfg, ref, cfg, eng = PU.make_pair(config, engine, preset, {}, "bce", 1e-4, 1.0, seed=seed)
    x, y = R.synth_batch(rcfg, 128, seed=10 + seed)
    res = PU.one_step_parity(f"{preset}_full_b128_t256_seed{seed}", rcfg, ref, cfg, eng, x, y, "bce", 1e-4, 1.0)
    assert res["perr"] <= PU.PROB_TOL


def test_c2_full_size_with_halo_boxes(pkg, monkeypatch):
    """The opt-in operand staging of conv_tc_kernel (SEDB200_CONV_HALO=1: one halo box per tap column serves the three
    tap rows; forward as fp16 + e4m3 passes on one tile, data gradients single-pass on tile pairs) against the same
    oracle step at the benchmarked size."""
    config, engine = pkg
    monkeypatch.setenv("SEDB200_CONV_HALO", "1")
    rcfg, ref, cfg, eng = PU.make_pair(config, engine, "c2", {}, "bce", 1e-4, 1.0, seed=3)
    x, y = R.synth_batch(rcfg, 128, seed=13)
    res = PU.one_step_parity("c2_full_b128_t256_halo", rcfg, ref, cfg, eng, x, y, "bce", 1e-4, 1.0)
    assert res["perr"] <= PU.PROB_TOL


def test_ragged_tile_pairs_in_the_conv_blocks(pkg):
    """seq_len 40 with 16 / 32 image rows per M tile: 3 / 2 tiles per image, so the last tile of an image is ragged and
    the last PAIR of tiles of an image holds one tile only -- at a batch where the single-pass launches do run on tile
    pairs (>= 3 pairs per CTA) and the halo boxes reach past the image on both sides."""
    config, engine = pkg
    rcfg, ref, cfg, eng = PU.make_pair(config, engine, "c2", {"seq_len": 40}, "bce", 1e-4, 1.0, seed=6)
    x, y = R.synth_batch(rcfg, 256, seed=17)
    res = PU.one_step_parity("c2_t40_b256_ragged_pairs", rcfg, ref, cfg, eng, x, y, "bce", 1e-4, 1.0)
    assert res["perr"] <= PU.PROB_TOL


def test_c5_long_context_t2048(pkg):
    """BASELINE configs[4] geometry at its real sequence length: T = 2048, 256 filters, 3 x BiGRU(128), 16 classes
    (batch 2: the recurrence length, not the batch, is what is new here -- 2,048 dependent fp32 steps per direction
    and layer, summed in a different order than MKL-DNN's GRU)."""
    config, engine = pkg
    rcfg, ref, cfg, eng = PU.make_pair(config, engine, "c5", {}, "bce", 1e-4, 1.0, seed=2)
    x, y = R.synth_batch(rcfg, 2, seed=21)
    res = PU.one_step_parity("c5_t2048_b2", rcfg, ref, cfg, eng, x, y, "bce", 1e-4, 1.0)
    assert res["perr"] <= PU.PROB_TOL


DROPOUT_CASES = [
    # preset, overrides, batch, tensor_cores
    ("c2", {"seq_len": 32}, 4, False),
    ("c2", {"seq_len": 32}, 4, True),
    ("c1", {"seq_len": 32}, 6, True),
    ("sedpy", {"seq_len": 32}, 4, True),            # fork layout, dropout after every block (sed.py:107)
    ("fork", {}, 16, False),                          # crnn_lightning.py:52: one dropout after the conv stack
    ("c2", {}, 128, True),                            # the benchmarked configuration itself
]


@pytest.mark.parametrize("preset,ov,batch,tc", DROPOUT_CASES)
def test_dropout_against_oracle_by_mask_injection(pkg, preset, ov, batch, tc):
    """Dropout ON (the benchmarked path).  The engine's masks come from a counter-based generator and are never
    stored; sedb200_crnn_dropout_mask exports them for (seed, block) and the oracle multiplies them in where
    nn.Dropout would draw its own (x * mask / (1 - p), sed.py:107).  Same one-step gates as with dropout off."""
    config, engine = pkg
    p = 0.4 if preset == "fork" else 0.5
    rcfg, ref, cfg, eng = PU.make_pair(config, engine, preset, ov, "bce", 1e-4, 1.0, seed=4, tensor_cores=tc,
                                       dropout=p, engine_seed=77)
    x, y = R.synth_batch(rcfg, batch, seed=31)
    exact = (not tc) and cfg.conv_ch > 64
    res = PU.one_step_parity(f"dropout_{preset}_{ov}_b{batch}_tc{int(tc)}", rcfg, ref, cfg, eng, x, y, "bce", 1e-4, 1.0,
                             masks_from_engine=True, exact_grads=exact)
    assert res["perr"] <= PU.PROB_TOL
    # the masks are real Bernoulli(1 - p) draws and differ between seeds / blocks
    m0 = eng.dropout_masks(batch, seed=1)
    m1 = eng.dropout_masks(batch, seed=2)
    last = len(m0) - 1
    frac = m0[last].float().mean().item()
    assert abs(frac - (1 - p)) < 0.02, frac
    assert not torch.equal(m0[last], m1[last])
    if cfg.dropout_each_block:
        assert not torch.equal(m0[0].flatten()[:4096], m0[1].flatten()[:4096])
    else:
        assert bool(m0[0].all())                      # blocks without dropout: all ones
