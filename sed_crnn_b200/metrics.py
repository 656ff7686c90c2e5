"""Segment-based metrics -- drop-in for /root/reference/metrics.py with the counting on the GPU.

Same function names and results as the reference (metrics.py:20-74): the integer counting
(TP / Nsys / Nref, per-row S / D / I, block maxima) runs in `sedb200_threshold_counts`; the final
float64 arithmetic below is the reference's, term for term, so results are bit-identical whenever the
thresholded decisions are.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

eps = np.finfo(float).eps          # utils.py:4

_COUNTS = {}


def _counts(O, T, block: int, threshold: float = 0.5) -> np.ndarray:
    """13 counts for decisions/probabilities O and references T ([N,T,C] or [rows,C]; numpy or torch)."""
    dev = O.device if isinstance(O, torch.Tensor) and O.is_cuda else torch.device("cuda")
    o = torch.as_tensor(np.ascontiguousarray(O) if isinstance(O, np.ndarray) else O).to(dev, torch.float32)
    t = torch.as_tensor(np.ascontiguousarray(T) if isinstance(T, np.ndarray) else T).to(dev, torch.float32)
    n_cls = o.shape[-1]
    o, t = o.reshape(-1, n_cls).contiguous(), t.reshape(-1, n_cls).contiguous()
    buf = _COUNTS.setdefault(dev, torch.zeros(13, dtype=torch.int64, device=dev))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sedb200_threshold_counts(o.data_ptr(), t.data_ptr(), o.shape[0], n_cls, int(block),
                                                       float(threshold), buf.data_ptr(), _lib.current_stream_ptr()))
    return buf.cpu().numpy()


def _f1(tp, nsys, nref):
    prec = float(tp) / float(nsys + eps)
    recall = float(tp) / float(nref + eps)
    return 2 * prec * recall / (prec + recall + eps)


def _er(s, d, i, nref):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.int64(s + d + i) / (np.float64(nref) + 0.0)


def scores_from_counts(c) -> tuple:
    """(f1_framewise, er_framewise, f1_1sec, er_1sec) from the 13 device counts."""
    c = [int(v) for v in c]
    return (_f1(c[0], c[1], c[2]), _er(c[3], c[4], c[5], c[2]),
            _f1(c[6], c[7], c[8]), _er(c[9], c[10], c[11], c[12]))


def f1_overall_framewise(O, T):
    return scores_from_counts(_counts(O, T, 1))[0]


def er_overall_framewise(O, T):
    return scores_from_counts(_counts(O, T, 1))[1]


def f1_overall_1sec(O, T, block_size):
    return scores_from_counts(_counts(O, T, block_size))[2]


def er_overall_1sec(O, T, block_size):
    return scores_from_counts(_counts(O, T, block_size))[3]


def compute_scores(pred, y, frames_in_1_sec=50):
    s = scores_from_counts(_counts(pred, y, frames_in_1_sec))
    return {"f1_overall_1sec": s[2], "er_overall_1sec": s[3]}
