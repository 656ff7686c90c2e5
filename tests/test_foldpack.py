"""Fold packing + the `.npz` `arr_0..arr_3` format and its consumers: feature.py:109-133 writes,
decorte_datamodule.py:24-34 / sed.py:115-125 read, decorte_datamodule.py:117-137 (DecorteDataModule) wraps.

CPU tests: the product reader against packs written by the oracle restatement of feature.py:109-133 (and, when
/root/reference is present, against the UNMODIFIED reference reader); the data-parallel batch split of the loader.
GPU tests: `feature.pack_folds` (device concat + device StandardScaler) against the restatement, file for file; the
DataModule end to end; `_spec_augment`."""
import os
import random

import numpy as np
import pytest

from oracle import dataset_ref


def test_reader_reads_reference_format(tmp_path):
    from sed_crnn_b200 import decorte_datamodule as DM
    vids = dataset_ref.synth_videos(6, 4, seed=3)
    paths = dataset_ref.pack_folds(vids, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == [f"mbe_mon_fold{i}.npz" for i in range(1, 5)]
    with np.load(paths[0]) as z:
        assert sorted(z.files) == ["arr_0", "arr_1", "arr_2", "arr_3"]
    got = DM._load_all_npz(str(tmp_path), verbose=False)
    want = dataset_ref.load_all_npz(str(tmp_path))
    assert sorted(got) == [1, 2, 3, 4]
    for i in want:
        for k in ("train_x", "train_y", "val_x", "val_y"):
            assert got[i][k].dtype == want[i][k].dtype and np.array_equal(got[i][k], want[i][k]), (i, k)
    # fold f's test split is exactly the videos assigned to fold f, standardised with the TRAIN statistics
    n_test = sum(v[0].shape[0] for v in vids.values() if v[2] == 0)
    assert got[1]["val_x"].shape == (n_test, 40) and got[1]["val_y"].shape == (n_test, 1)
    assert abs(float(got[1]["train_x"].mean())) < 1e-5


@pytest.mark.reference
def test_reader_matches_unmodified_reference_reader(tmp_path, capsys):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("/root/reference not present")
    from sed_crnn_b200 import decorte_datamodule as DM
    dataset_ref.pack_folds(dataset_ref.synth_videos(5, 4, seed=1), str(tmp_path))
    ref = ref_import.load("decorte_datamodule")._load_all_npz(str(tmp_path))
    got = DM._load_all_npz(str(tmp_path), verbose=False)
    for i in ref:
        for k in ref[i]:
            assert np.array_equal(got[i][k], ref[i][k])


class _FakeDS:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def batch(self, idx):
        return list(idx)


@pytest.mark.parametrize("n,bs,world,drop_last", [(1000, 128, 2, False), (1000, 128, 8, False), (1025, 128, 8, False),
                                                  (1000, 128, 4, True), (130, 128, 8, False)])
def test_loader_gives_every_rank_the_same_number_of_batches(n, bs, world, drop_last):
    """One gradient exchange per batch: ranks must agree on the batch count, also for the last partial global batch
    (ranks with an empty slice used to skip it and the all-reduce hung)."""
    from sed_crnn_b200.decorte_datamodule import DeviceWindowLoader
    per_rank = [list(DeviceWindowLoader(_FakeDS(n), bs, shuffle=True, drop_last=drop_last, rank=r, world_size=world,
                                        seed=5)) for r in range(world)]
    counts = {len(b) for b in per_rank}
    assert len(counts) == 1 and counts.pop() == (n // bs if drop_last else -(-n // bs))
    assert all(len(batch) >= 1 for b in per_rank for batch in b)
    for step in range(len(per_rank[0]) - (0 if drop_last else 1)):            # full global batches: a partition
        idx = sorted(i for r in range(world) for i in per_rank[r][step])
        assert len(idx) == bs and len(set(idx)) == bs
    seen = {i for b in per_rank for batch in b for i in batch}
    assert seen == set(range(n)) if not drop_last else len(seen) == (n // bs) * bs
    # next epoch: another permutation, still identical across ranks
    l0 = DeviceWindowLoader(_FakeDS(n), bs, shuffle=True, rank=0, world_size=world, seed=5)
    e0, e1 = list(l0), list(l0)
    assert e0 != e1


def test_loader_refuses_unseeded_shuffle_across_ranks():
    from sed_crnn_b200.decorte_datamodule import DeviceWindowLoader
    with pytest.raises(ValueError):
        DeviceWindowLoader(_FakeDS(10), 4, shuffle=True, rank=0, world_size=2)


# ----------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_pack_folds_matches_reference_restatement(built_lib, tmp_path):
    import torch
    from sed_crnn_b200 import feature
    vids = dataset_ref.synth_videos(7, 4, seed=11)
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    # half of the videos arrive as device tensors (straight from mbe_device), half as numpy (the per-video cache)
    mixed = {k: ((torch.from_numpy(m).cuda(), torch.from_numpy(l).cuda(), f) if i % 2 else (m, l, f))
             for i, (k, (m, l, f)) in enumerate(vids.items())}
    ours = feature.pack_folds(mixed, str(a))
    ref = dataset_ref.pack_folds(vids, str(b))
    assert [os.path.basename(p) for p in ours] == [os.path.basename(p) for p in ref]
    for po, pr in zip(ours, ref):
        with np.load(po) as zo, np.load(pr) as zr:
            assert zo.files == zr.files == ["arr_0", "arr_1", "arr_2", "arr_3"]
            for k in zr.files:
                assert zo[k].dtype == zr[k].dtype == np.float32 and zo[k].shape == zr[k].shape, k
            assert np.array_equal(zo["arr_1"], zr["arr_1"]) and np.array_equal(zo["arr_3"], zr["arr_3"])
            for k in ("arr_0", "arr_2"):
                assert np.abs(zo[k] - zr[k]).max() <= 5e-6, (k, np.abs(zo[k] - zr[k]).max())


@pytest.mark.gpu
def test_datamodule_end_to_end(built_lib, tmp_path):
    import torch
    from sed_crnn_b200 import decorte_datamodule as DM
    dataset_ref.pack_folds(dataset_ref.synth_videos(6, 4, seed=2), str(tmp_path))
    dm = DM.DecorteDataModule(fold_id=2, cache_dir=str(tmp_path), batch_size=16, num_workers=4)
    dm.setup()
    ref = dataset_ref.load_all_npz(str(tmp_path))[2]
    assert dm.train_ds.augment and not dm.val_ds.augment
    assert dm.val_ds.total_frames == ref["val_x"].shape[0]
    # validation loader: no shuffle, no augmentation -> items replay exactly from the host RNG streams
    random.seed(123)
    got = [(x.cpu().numpy(), y.cpu().numpy()) for x, y in dm.val_dataloader()]
    assert len(got) == len(dm.val_dataloader()) and got[0][0].shape == (16, 1, 40, 64) and got[0][1].shape == (16, 8, 1)
    random.seed(123)
    pos = np.where(ref["val_y"][:, 0] == 1)[0].tolist()
    neg = dataset_ref.find_clean_negatives(ref["val_y"], 64).tolist()
    draws = DM.WindowDraws(pos, neg, ref["val_x"].shape[0], 40, False, 64)
    k = 0
    for xb, yb in got:
        for j in range(xb.shape[0]):
            start, _, _ = draws.draw(k)
            x, y = dataset_ref.window_item(ref["val_x"], ref["val_y"], start, 64, 8)
            assert np.array_equal(xb[j], x) and np.array_equal(yb[j], y), k
            k += 1
    assert k == len(dm.val_ds)
    # training loader: shuffled, drop_last, augmented, CUDA tensors of the reference's batch shapes
    tl = dm.train_dataloader()
    xb, yb = next(iter(tl))
    assert xb.is_cuda and xb.shape == (16, 1, 40, 64) and yb.shape == (16, 8, 1)
    assert len(tl) == len(dm.train_ds) // 16
    assert bool((xb == 0).any())                         # SpecAugment zeroed something


@pytest.mark.gpu
def test_spec_augment_matches_reference_restatement(built_lib):
    import torch
    from sed_crnn_b200 import decorte_datamodule as DM
    rng = np.random.default_rng(0)
    mel = rng.standard_normal((40, 64)).astype(np.float32)
    np.random.seed(9)
    want = dataset_ref.spec_augment(mel.copy())
    np.random.seed(9)
    got = DM._spec_augment(torch.from_numpy(mel).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    np.random.seed(9)
    arr = mel.copy()
    out = DM._spec_augment(arr)
    assert out is arr and np.array_equal(arr, want)
