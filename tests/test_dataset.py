"""Window sampler + label rasteriser: reference decorte_datamodule.py:19-111, feature.py:88-93.

CPU tests pin the oracle restatement and the host-side RNG replay against tests/golden/window_sampler.npz (items of
the UNMODIFIED reference HitWindowDataset); GPU tests compare the kernels bit-exactly with the golden items and
the oracle."""
import os
import random

import numpy as np
import pytest

from oracle import dataset_ref


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "window_sampler.npz"))


def _replay(gold, aug):
    """Draw the 24 golden items' (start, masks) with the host mirror under the golden seeds."""
    from sed_crnn_b200.decorte_datamodule import WindowDraws
    lab = gold["lab"]
    draws = WindowDraws(np.where(lab[:, 0] == 1)[0].tolist(), dataset_ref.find_clean_negatives(lab, 64).tolist(),
                        lab.shape[0], 40, aug, 64)
    random.seed(int(gold["seeds"][0]))
    np.random.seed(int(gold["seeds"][1]))
    return [draws.draw(int(i)) for i in gold["idx"]]


def test_oracle_clean_negatives_matches_reference(gold):
    assert np.array_equal(dataset_ref.find_clean_negatives(gold["lab"], 64), gold["neg_starts"])


@pytest.mark.parametrize("name,aug", [("plain", False), ("aug", True)])
def test_oracle_and_rng_replay_match_reference_items(gold, name, aug):
    draws = _replay(gold, aug)
    for i, (start, t0, f0) in enumerate(draws):
        x, y = dataset_ref.window_item(gold["mel"], gold["lab"], start, 64, 8, t0, f0)
        assert np.array_equal(x, gold[f"{name}_x"][i]), f"item {i}"
        assert np.array_equal(y, gold[f"{name}_y"][i]), f"item {i}"


def test_oracle_rasterize_python_slice_semantics():
    lbl = dataset_ref.rasterize_labels([0.0, 1.0, 2.5], [0.01, 1.5, 400.0], 120)
    assert lbl[0, 0] == 1 and lbl[1, 0] == 0
    assert lbl[43:65, 0].all() and lbl[42, 0] == 0 and lbl[65, 0] == 0          # floor(43.07) .. ceil(64.6)
    assert lbl[107:, 0].all()                                                    # end clamps to n_frames
    assert lbl.sum() == 1 + 22 + (120 - 107)


@pytest.mark.reference
def test_golden_is_current_reference_output(gold):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("/root/reference not present")
    dm = ref_import.load("decorte_datamodule")
    random.seed(int(gold["seeds"][0]))
    np.random.seed(int(gold["seeds"][1]))
    ds = dm.HitWindowDataset(gold["mel"], gold["lab"], augment=True)
    x0, y0 = ds[0]
    assert np.array_equal(x0.numpy(), gold["aug_x"][0]) and np.array_equal(y0.numpy(), gold["aug_y"][0])
    assert len(ds) == int(gold["aug_len"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_clean_negatives_kernel(gold, built_lib):
    from sed_crnn_b200 import decorte_datamodule as dm
    assert np.array_equal(dm._find_clean_negatives(gold["lab"]), gold["neg_starts"])
    assert dm._find_clean_negatives(np.zeros((10, 1), np.float32)).size == 0       # shorter than one window
    lab = np.zeros((64, 1), np.float32)
    assert np.array_equal(dm._find_clean_negatives(lab), [0])                      # exactly one window
    lab[63] = 1
    assert dm._find_clean_negatives(lab).size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,aug", [("plain", False), ("aug", True)])
def test_hit_window_dataset_bit_exact_vs_reference(gold, built_lib, name, aug):
    from sed_crnn_b200 import decorte_datamodule as dm
    ds = dm.HitWindowDataset(gold["mel"], gold["lab"], augment=aug)
    assert len(ds) == int(gold[f"{name}_len"])
    assert np.array_equal(ds.pos_frames, gold[f"{name}_pos_frames"])
    random.seed(int(gold["seeds"][0]))
    np.random.seed(int(gold["seeds"][1]))
    x, y = ds.batch(gold["idx"])
    assert x.shape == (24, 1, 40, 64) and y.shape == (24, 8, 1)
    assert np.array_equal(x.cpu().numpy(), gold[f"{name}_x"])
    assert np.array_equal(y.cpu().numpy(), gold[f"{name}_y"])
    random.seed(int(gold["seeds"][0]))
    np.random.seed(int(gold["seeds"][1]))
    x0, y0 = ds[0]                                                                  # per-item form
    assert np.array_equal(x0.cpu().numpy(), gold[f"{name}_x"][0]) and tuple(y0.shape) == (8, 1)


@pytest.mark.gpu
def test_window_batch_ragged_shapes_and_sednet_layout(built_lib):
    import torch
    from sed_crnn_b200 import decorte_datamodule as dm
    rng = np.random.default_rng(3)
    n, n_ch, F, K, L = 1500, 2, 40, 6, 256
    mel = rng.standard_normal((n, n_ch * F)).astype(np.float32)
    lab = (rng.random((n, K)) < 0.2).astype(np.float32)
    starts = np.concatenate([[0, n - L], rng.integers(0, n - L + 1, 35)])
    d_mel, d_lab = torch.from_numpy(mel).cuda(), torch.from_numpy(lab).cuda()
    x, y = dm.window_batch(d_mel, d_lab, starts, seq_in=L, seq_out=L, n_ch=n_ch, layout=dm.LAYOUT_SEDNET)
    wx, wy = dataset_ref.window_batch_sednet(mel, lab, starts, L, n_ch)
    assert np.array_equal(x.cpu().numpy(), wx) and np.array_equal(y.cpu().numpy(), wy)
    # fork layout, odd sizes (F and L not multiples of the 32 x 32 tile), masks incl. disabled (-1) entries
    F2, L2 = 37, 50
    mel2 = rng.standard_normal((300, F2)).astype(np.float32)
    lab2 = (rng.random((300, 1)) < 0.3).astype(np.float32)
    st2 = rng.integers(0, 300 - L2 + 1, 9)
    tm = rng.integers(-1, L2 - 8, (9, 3))
    fm = rng.integers(-1, F2 - 8, (9, 3))
    x2, y2 = dm.window_batch(torch.from_numpy(mel2).cuda(), torch.from_numpy(lab2).cuda(), st2, seq_in=L2, seq_out=10,
                             tmask=tm, fmask=fm)
    for i in range(9):
        wx2, wy2 = dataset_ref.window_item(mel2, lab2, int(st2[i]), L2, 10, tm[i], fm[i])
        assert np.array_equal(x2[i].cpu().numpy(), wx2) and np.array_equal(y2[i].cpu().numpy(), wy2)
    # empty batch is a no-op, too-short matrix is an error
    xe, ye = dm.window_batch(d_mel, d_lab, [], seq_in=L, seq_out=L, n_ch=n_ch)
    assert xe.shape[0] == 0
    with pytest.raises(RuntimeError):
        dm.window_batch(d_mel[:100].contiguous(), None, [0], seq_in=L, seq_out=L, n_ch=n_ch)


@pytest.mark.gpu
def test_device_window_loader_epoch(gold, built_lib):
    import torch
    from sed_crnn_b200 import decorte_datamodule as dm
    ds = dm.HitWindowDataset(gold["mel"], gold["lab"], augment=True)
    loader = dm.DeviceWindowLoader(ds, batch_size=32, shuffle=True, drop_last=True,
                                   generator=torch.Generator().manual_seed(0))
    assert len(loader) == len(ds) // 32
    nb = 0
    for xb, yb in loader:
        assert xb.is_cuda and xb.shape == (32, 1, 40, 64) and yb.shape == (32, 8, 1)
        assert set(np.unique(yb.cpu().numpy())) <= {0.0, 1.0}
        nb += 1
    assert nb == len(loader)
    halves = [dm.DeviceWindowLoader(ds, 32, rank=r, world_size=2) for r in (0, 1)]
    assert all(next(iter(h))[0].shape[0] == 16 for h in halves)


@pytest.mark.gpu
def test_rasterize_labels_kernel(built_lib):
    from sed_crnn_b200 import feature
    rng = np.random.default_rng(9)
    n = 7752
    a = np.sort(rng.random(40) * 185.0)
    b = a + rng.random(40) * 2.0
    a[0], b[0] = 0.0, 0.0                                   # empty event
    a[1], b[1] = 1024 / 44100, 2048 / 44100                 # exact frame boundaries
    got = feature.rasterize_labels(a, b, n).cpu().numpy()
    assert np.array_equal(got, dataset_ref.rasterize_labels(a, b, n))
    assert feature.rasterize_labels([], [], 10).sum().item() == 0
