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
