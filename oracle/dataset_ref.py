"""CPU restatement of the reference's window sampler and label rasteriser.  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg, never by the product.

Pinned by tests/golden/window_sampler.npz, which oracle/make_golden.py produces by running the UNMODIFIED
`decorte_datamodule.HitWindowDataset` / `_find_clean_negatives` / `_spec_augment` of /root/reference
(pytorch_lightning stubbed, it does no arithmetic).  The rasteriser (feature.py:89-93) lives inside the reference's
`__main__` block and cannot be imported; it is restated line by line below.
"""
from __future__ import annotations

import numpy as np


def find_clean_negatives(label_vec: np.ndarray, seq_len_in: int) -> np.ndarray:
    """decorte_datamodule.py:19-23 / sed.py:48-52."""
    mask = (label_vec[:, 0] == 1).astype(np.uint8)
    window = np.ones(seq_len_in, dtype=np.uint8)
    overlap = np.convolve(mask, window, mode="valid")
    return np.where(overlap == 0)[0]


def window_item(mel: np.ndarray, lab: np.ndarray, start: int, seq_in: int, seq_out: int, t0=(), f0=(),
                time_mask_w: int = 8, freq_mask_w: int = 8):
    """decorte_datamodule.py:96-111 for one item with the random draws given: x (1, n_mel, seq_in), y (seq_out, 1)."""
    x = mel[start:start + seq_in].T.copy()                                   # :96
    for a, b in zip(t0, f0):                                                  # :40-48
        if a >= 0:
            x[:, a:a + time_mask_w] = 0.0
        if b >= 0:
            x[b:b + freq_mask_w, :] = 0.0
    lab_win = lab[start:start + seq_in]
    y = lab_win.reshape(seq_out, -1).max(axis=1, keepdims=True)               # :101
    return x[None].astype(np.float32), y.astype(np.float32)


def window_batch_sednet(mel: np.ndarray, lab: np.ndarray, starts, seq_in: int, n_ch: int):
    """Time-major multi-channel windows (SEDnet layout, utils.split_multi_channels order): x [B, n_ch, seq_in, F],
    y [B, seq_in, K]."""
    F = mel.shape[1] // n_ch
    x = np.stack([mel[s:s + seq_in].reshape(seq_in, n_ch, F).transpose(1, 0, 2) for s in starts])
    y = np.stack([lab[s:s + seq_in] for s in starts])
    return x.astype(np.float32), y.astype(np.float32)


def rasterize_labels(starts_s, ends_s, n_frames: int, sr: int = 44100, hop: int = 1024) -> np.ndarray:
    """feature.py:88-93."""
    lbl = np.zeros((n_frames, 1), dtype=np.float32)
    for a, b in zip(starts_s, ends_s):
        s = int(np.floor(a * sr / hop))
        e = int(np.ceil(b * sr / hop))
        lbl[s:e, 0] = 1.0
    return lbl


def spec_augment(mel: np.ndarray, masks_per_ex: int = 2, time_mask_w: int = 8, freq_mask_w: int = 8) -> np.ndarray:
    """decorte_datamodule.py:39-49 (in place, draws from np.random in the reference's order)."""
    for _ in range(masks_per_ex):
        if mel.shape[1] > time_mask_w:
            t0 = np.random.randint(0, mel.shape[1] - time_mask_w)
            mel[:, t0:t0 + time_mask_w] = 0.0
        if mel.shape[0] > freq_mask_w:
            f0 = np.random.randint(0, mel.shape[0] - freq_mask_w)
            mel[f0:f0 + freq_mask_w, :] = 0.0
    return mel


def pack_folds(per_video: dict, cache_dir: str) -> list:
    """feature.py:109-133, line by line (sklearn.preprocessing.StandardScaler is the reference's own scaler)."""
    import os
    from sklearn import preprocessing
    paths = []
    fold_k = max(v[2] for v in per_video.values()) + 1                        # :112
    for f in range(fold_k):
        X_train, Y_train, X_test, Y_test = None, None, None, None
        for vname, (mbe, lbl, fold) in per_video.items():                     # :116-122
            if fold == f:
                X_test = mbe if X_test is None else np.concatenate((X_test, mbe), axis=0)
                Y_test = lbl if Y_test is None else np.concatenate((Y_test, lbl), axis=0)
            else:
                X_train = mbe if X_train is None else np.concatenate((X_train, mbe), axis=0)
                Y_train = lbl if Y_train is None else np.concatenate((Y_train, lbl), axis=0)
        scaler = preprocessing.StandardScaler()                               # :126-128
        X_train = scaler.fit_transform(X_train)
        X_test = scaler.transform(X_test)
        out_fold = os.path.join(cache_dir, f"mbe_mon_fold{f+1}.npz")          # :130-131
        np.savez(out_fold, X_train, Y_train, X_test, Y_test)
        paths.append(out_fold)
    return paths


def load_all_npz(folder: str) -> dict:
    """decorte_datamodule.py:24-34 / sed.py:115-125."""
    import os
    folds = {}
    for i in range(1, 5):
        arr = np.load(os.path.join(folder, f"mbe_mon_fold{i}.npz"))
        folds[i] = {"train_x": arr["arr_0"], "train_y": arr["arr_1"], "val_x": arr["arr_2"], "val_y": arr["arr_3"]}
    return folds


def synth_videos(n_videos: int = 6, n_folds: int = 4, seed: int = 0, n_feat: int = 40) -> dict:
    """Seeded stand-in for the per-video cache of feature.py:70-107: {name: (mbe, lbl, fold_id)} with round-robin
    folds (decorte_data_loader.py assigns folds round-robin) and a few positive label runs per video."""
    rng = np.random.default_rng(seed)
    out = {}
    for v in range(n_videos):
        frames = int(rng.integers(150, 260))
        mbe = (rng.standard_normal((frames, n_feat)) * rng.uniform(0.5, 3.0, n_feat) + rng.uniform(-8, 2, n_feat)).astype(np.float32)
        lbl = np.zeros((frames, 1), dtype=np.float32)
        for _ in range(4):
            a = int(rng.integers(0, frames - 6))
            lbl[a:a + int(rng.integers(1, 6)), 0] = 1.0
        out[f"video_{v:02d}.mp4"] = (mbe, lbl, v % n_folds)
    return out
