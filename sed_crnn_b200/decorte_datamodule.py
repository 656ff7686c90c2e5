"""Drop-in for the window sampler of /root/reference/decorte_datamodule.py (and sed.py:48-79) with the per-item
numpy slicing moved onto the GPU.

Same names as the reference: `_find_clean_negatives`, `HitWindowDataset(mel, lab, augment)` with `pos_frames`,
`neg_starts`, `total_frames`, `__len__`, `_rand_pos`, `_rand_neg`, `__getitem__`.  The fold's feature / label
matrices are uploaded ONCE and stay in HBM; a batch is one kernel launch (`sedb200_window_batch_f32`) driven by
the window starts and SpecAugment offsets, which are drawn on the host with exactly the reference's RNG calls in
the reference's order (`random.choice` / `random.randint` for the start, `np.random.randint` for each mask,
decorte_datamodule.py:41-48,67-74,89) -- so under the same seeds the windows are bit-identical to the reference's.

`DeviceWindowLoader` replaces `DataLoader(HitWindowDataset, batch_size, shuffle, drop_last)`
(decorte_datamodule.py:127-137): it yields CUDA tensors `[B,1,N_MELS,SEQ_LEN_IN]`, `[B,SEQ_LEN_OUT,1]`.
`DecorteDataModule(fold_id, cache_dir, batch_size, num_workers)` (decorte_datamodule.py:117-137), `_load_all_npz`
(:24-34) and `_spec_augment` (:39-49) keep the reference names, so `from decorte_datamodule import DecorteDataModule`
(train_lightning.py:12) works against this module.  `HitWindowDataset.__getitem__` returns CUDA tensors, so it must
NOT be wrapped in a `DataLoader(num_workers > 0, pin_memory=True)` -- use the loaders the DataModule returns.
There is no CPU path: every window is produced by the kernel.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import _lib
from .train_constants import (BATCH_SIZE, FREQ_MASK_W, MASKS_PER_EX, NUM_WORKERS, SEQ_LEN_IN, SEQ_LEN_OUT,
                              TIME_MASK_W)

CACHE_DIR = os.path.expanduser("~/src/plai_cv/cache/decorte_metadata/features")     # train_constants.py:9

LAYOUT_FORK, LAYOUT_SEDNET = 0, 1


def _as_dev(a, dtype=torch.float32) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            a = a.cuda()
        return a.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype).contiguous()


def _find_clean_negatives(label_vec, seq_len_in: int = SEQ_LEN_IN) -> np.ndarray:
    """decorte_datamodule.py:19-23: starts whose SEQ_LEN_IN-frame window holds no positive frame (sorted int64)."""
    lab = _as_dev(label_vec)
    if lab.dim() == 1:
        lab = lab[:, None].contiguous()
    n = lab.shape[0] - seq_len_in + 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    flag = torch.empty(n, dtype=torch.uint8, device=lab.device)
    with torch.cuda.device(lab.device):
        _lib.check(_lib.lib().sedb200_clean_negatives(lab.data_ptr(), lab.shape[0], lab.shape[1], seq_len_in,
                                                      flag.data_ptr(), _lib.current_stream_ptr()))
    return np.flatnonzero(flag.cpu().numpy()).astype(np.int64)


def _load_all_npz(folder: str, folds=range(1, 5), verbose: bool = True) -> dict:
    """decorte_datamodule.py:24-34 / sed.py:115-125: read the fold packs `mbe_mon_fold{i}.npz` (keys `arr_0..arr_3` =
    train_x, train_y, val_x, val_y, written by feature.py:131 / `feature.pack_folds`) into host RAM."""
    out = {}
    for i in folds:
        fp = os.path.join(folder, f"mbe_mon_fold{i}.npz")
        arr = np.load(fp)
        out[i] = {"train_x": arr["arr_0"], "train_y": arr["arr_1"], "val_x": arr["arr_2"], "val_y": arr["arr_3"]}
        if verbose:
            print(f"loaded into RAM -> fold {i}  ({arr['arr_0'].nbytes / 1e6:0.1f} MB train)")
    return out


def _spec_augment(mel):
    """decorte_datamodule.py:39-49 for ONE window `mel` [N_MELS, SEQ_LEN_IN] (CUDA tensor or numpy array): draws the
    mask offsets from `np.random` exactly like the reference (time offset, then frequency offset, MASKS_PER_EX times)
    and zeroes the bands with the window kernel's masking path.  Returns the same kind of object it was given."""
    was_numpy = not isinstance(mel, torch.Tensor)
    m = _as_dev(mel)
    F, T = m.shape
    t0, f0 = [-1] * MASKS_PER_EX, [-1] * MASKS_PER_EX
    for i in range(MASKS_PER_EX):
        if T > TIME_MASK_W:
            t0[i] = int(np.random.randint(0, T - TIME_MASK_W))
        if F > FREQ_MASK_W:
            f0[i] = int(np.random.randint(0, F - FREQ_MASK_W))
    x, _ = window_batch(m.t().contiguous(), None, [0], seq_in=T, tmask=[t0], fmask=[f0])
    out = x[0, 0]
    if was_numpy:
        mel[...] = out.cpu().numpy()             # the reference masks in place and returns its argument
        return mel
    return out


def window_batch(mel: torch.Tensor, lab: torch.Tensor | None, starts, *, seq_in: int = SEQ_LEN_IN,
                 seq_out: int = SEQ_LEN_OUT, n_ch: int = 1, tmask=None, fmask=None, time_mask_w: int = TIME_MASK_W,
                 freq_mask_w: int = FREQ_MASK_W, layout: int = LAYOUT_FORK, out_x: torch.Tensor | None = None,
                 out_y: torch.Tensor | None = None):
    """One launch: windows `mel[s:s+seq_in]` for every start -> x, pooled labels -> y (see include/sedb200.h)."""
    if not (isinstance(mel, torch.Tensor) and mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous()):
        raise TypeError("mel must be a contiguous CUDA float32 tensor [frames, n_ch*n_feat] (no CPU fallback)")
    dev = mel.device
    n_frames, cols = mel.shape
    if cols % n_ch:
        raise ValueError(f"{cols} feature columns do not split into {n_ch} channels")
    F = cols // n_ch
    st = torch.as_tensor(np.asarray(starts, dtype=np.int64)) if not isinstance(starts, torch.Tensor) else starts
    st = st.to(dev, torch.int64).contiguous()
    B = st.numel()
    n_masks = 0
    tm = fm = None
    if tmask is not None or fmask is not None:
        ref = tmask if tmask is not None else fmask
        n_masks = int(np.asarray(ref).shape[-1]) if not isinstance(ref, torch.Tensor) else ref.shape[-1]
        if tmask is not None:
            tm = torch.as_tensor(np.asarray(tmask, dtype=np.int32)).to(dev).reshape(B, n_masks).contiguous()
        if fmask is not None:
            fm = torch.as_tensor(np.asarray(fmask, dtype=np.int32)).to(dev).reshape(B, n_masks).contiguous()
    xshape = (B, n_ch, F, seq_in) if layout == LAYOUT_FORK else (B, n_ch, seq_in, F)
    x = out_x if out_x is not None else torch.empty(xshape, dtype=torch.float32, device=dev)
    if tuple(x.shape) != xshape or not x.is_contiguous():
        raise ValueError(f"out_x must be contiguous {xshape}")
    y = None
    K = 0
    if lab is not None:
        if not (lab.is_cuda and lab.dtype == torch.float32 and lab.is_contiguous() and lab.shape[0] == n_frames):
            raise TypeError("lab must be a contiguous CUDA float32 tensor [frames, n_lab]")
        K = lab.shape[1]
        y = out_y if out_y is not None else torch.empty(B, seq_out, K, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sedb200_window_batch_f32(
            mel.data_ptr(), lab.data_ptr() if lab is not None else None, n_frames, n_ch, F, K, st.data_ptr(), B,
            seq_in, seq_out, tm.data_ptr() if tm is not None else None, fm.data_ptr() if fm is not None else None,
            n_masks, time_mask_w, freq_mask_w, layout, x.data_ptr(), y.data_ptr() if y is not None else None,
            _lib.current_stream_ptr()))
    return x, y


class WindowDraws:
    """Host half of `HitWindowDataset`: which window / which masks, consuming the RNG streams exactly like the
    reference (`random` for the start, `np.random` for the masks; decorte_datamodule.py:41-48,67-74,89-94)."""

    def __init__(self, pos_frames, neg_starts, total_frames: int, n_mel: int, augment: bool,
                 seq_len_in: int = SEQ_LEN_IN):
        self.pos_frames, self.neg_starts = list(pos_frames), list(neg_starts)
        self.total_frames, self.n_mel, self.augment, self.seq_len_in = total_frames, n_mel, augment, seq_len_in

    def _rand_pos(self):
        center = random.choice(self.pos_frames)
        a = max(0, center - self.seq_len_in + 1)
        b = min(center, self.total_frames - self.seq_len_in)
        return random.randint(a, b)

    def _rand_neg(self):
        return random.choice(self.neg_starts)

    def draw(self, idx: int):
        start = self._rand_pos() if idx % 2 == 0 else self._rand_neg()
        if start + self.seq_len_in > self.total_frames:                       # decorte_datamodule.py:92-94
            start = max(0, self.total_frames - self.seq_len_in)
        t0 = [-1] * MASKS_PER_EX
        f0 = [-1] * MASKS_PER_EX
        if self.augment:
            for i in range(MASKS_PER_EX):                                     # decorte_datamodule.py:40-48
                if self.seq_len_in > TIME_MASK_W:
                    t0[i] = int(np.random.randint(0, self.seq_len_in - TIME_MASK_W))
                if self.n_mel > FREQ_MASK_W:
                    f0[i] = int(np.random.randint(0, self.n_mel - FREQ_MASK_W))
        return start, t0, f0


class HitWindowDataset:
    """decorte_datamodule.py:54-111 / sed.py:55-76 with device-resident matrices.  `__getitem__` returns the same
    `(x[1,N_MELS,SEQ_LEN_IN], y[SEQ_LEN_OUT,1])` pair (as CUDA tensors); `batch(indices)` is the fast path."""

    def __init__(self, mel, lab, augment: bool = False, seq_len_in: int = SEQ_LEN_IN, seq_len_out: int = SEQ_LEN_OUT):
        lab_np = lab.cpu().numpy() if isinstance(lab, torch.Tensor) else np.asarray(lab)
        if lab_np.ndim == 1:
            lab_np = lab_np[:, None]
        self.mel, self.lab = _as_dev(mel), _as_dev(lab_np)
        self.augment = augment
        self.seq_len_in, self.seq_len_out = seq_len_in, seq_len_out
        self.total_frames = self.mel.shape[0]
        self._draws = WindowDraws(np.where(lab_np[:, 0] == 1)[0].tolist(),
                                  _find_clean_negatives(self.lab, seq_len_in).tolist(), self.total_frames,
                                  self.mel.shape[1], augment, seq_len_in)
        self.pos_frames, self.neg_starts = self._draws.pos_frames, self._draws.neg_starts

    def __len__(self):
        return len(self.pos_frames) * 2

    def _rand_pos(self):
        return self._draws._rand_pos()

    def _rand_neg(self):
        return self._draws._rand_neg()

    def batch(self, indices):
        draws = [self._draws.draw(int(i)) for i in indices]
        starts = [d[0] for d in draws]
        tm = [d[1] for d in draws] if self.augment else None
        fm = [d[2] for d in draws] if self.augment else None
        return window_batch(self.mel, self.lab, starts, seq_in=self.seq_len_in, seq_out=self.seq_len_out,
                            tmask=tm, fmask=fm)

    def __getitem__(self, idx):
        x, y = self.batch([idx])
        return x[0], y[0]


class DeviceWindowLoader:
    """`DataLoader(ds, batch_size, shuffle=..., drop_last=...)` (decorte_datamodule.py:127-137) for a device-resident
    `HitWindowDataset`: one kernel launch per batch, no worker processes.  With `world_size > 1` every rank takes
    its contiguous share of each global batch (parallel.batch_slice_for_rank semantics)."""

    def __init__(self, ds: HitWindowDataset, batch_size: int = BATCH_SIZE, shuffle: bool = False,
                 drop_last: bool = False, generator: torch.Generator | None = None, rank: int = 0,
                 world_size: int = 1, seed: int | None = None):
        self.ds, self.batch_size, self.shuffle, self.drop_last = ds, batch_size, shuffle, drop_last
        self.generator, self.rank, self.world_size = generator, rank, world_size
        self.seed, self.epoch = seed, 0
        if batch_size % world_size:
            raise ValueError(f"global batch {batch_size} does not split over {world_size} ranks")
        if world_size > 1 and shuffle and generator is None and seed is None:
            # every rank must walk the SAME permutation, or "rank r's slice of global batch b" means nothing
            raise ValueError("DeviceWindowLoader(world_size > 1, shuffle=True) needs `seed` (shared by all ranks) or "
                             "an identically seeded `generator`")

    def __len__(self):
        n = len(self.ds)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def __iter__(self):
        n = len(self.ds)
        gen = self.generator
        if self.shuffle and gen is None and self.seed is not None:
            gen = torch.Generator().manual_seed(self.seed + self.epoch)
        order = torch.randperm(n, generator=gen).tolist() if self.shuffle else list(range(n))
        self.epoch += 1
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            if self.world_size > 1:
                # every rank must yield the SAME number of batches (one gradient exchange per batch): the last, partial
                # global batch is split as evenly as its size allows, and when it holds fewer items than ranks the
                # ranks left over re-use its first items (their gradients enter the average like anyone's)
                from .parallel import shard_range
                r = shard_range(len(idx), self.rank, self.world_size)
                idx = idx[r.start:r.stop] if len(r) else idx[:1]
            yield self.ds.batch(idx)


try:                                                       # real Lightning when it is installed
    import pytorch_lightning as pl
    _DataModuleBase = pl.LightningDataModule
except Exception:                                          # pragma: no cover - not installed in this image
    class _DataModuleBase:
        def __init__(self):
            pass


class DecorteDataModule(_DataModuleBase):
    """Drop-in for decorte_datamodule.DecorteDataModule (decorte_datamodule.py:117-137): same constructor, `setup`,
    `train_dataloader`, `val_dataloader`, attributes `train_ds` / `val_ds`.  The fold pack is read from
    `cache_dir/mbe_mon_fold{fold_id}.npz` and uploaded once; the loaders are `DeviceWindowLoader`s (one kernel launch
    per batch) instead of `DataLoader(num_workers=4, pin_memory=True)` -- `num_workers` is accepted and ignored.
    `rank` / `world_size` / `seed` shard every global batch over data-parallel ranks (reference: devices=1)."""

    def __init__(self, fold_id: int, cache_dir: str = CACHE_DIR, batch_size: int = BATCH_SIZE,
                 num_workers: int = NUM_WORKERS, rank: int = 0, world_size: int = 1, seed: int | None = None):
        super().__init__()
        self.fold_id, self.cache_dir = fold_id, cache_dir
        self.batch_size, self.num_workers = batch_size, num_workers
        self.rank, self.world_size, self.seed = rank, world_size, seed

    def setup(self, stage=None):
        fdata = _load_all_npz(self.cache_dir, folds=[self.fold_id])[self.fold_id]    # the reference loads all four
        self.train_ds = HitWindowDataset(fdata["train_x"], fdata["train_y"], augment=True)
        self.val_ds = HitWindowDataset(fdata["val_x"], fdata["val_y"], augment=False)

    def train_dataloader(self):
        return DeviceWindowLoader(self.train_ds, self.batch_size, shuffle=True, drop_last=True, rank=self.rank,
                                  world_size=self.world_size, seed=self.seed)

    def val_dataloader(self):
        return DeviceWindowLoader(self.val_ds, self.batch_size, shuffle=False, rank=self.rank,
                                  world_size=self.world_size)
