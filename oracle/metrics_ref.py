"""CPU restatement of the reference's segment metrics.  TEST INFRASTRUCTURE (see oracle/__init__).

Follows /root/reference/metrics.py:14-74 and /root/reference/utils.py:4,11-12:

  f1_overall_framewise  metrics.py:20-29   TP = #(T==1 & O==1); prec = TP/(Nsys+eps) ...
  er_overall_framewise  metrics.py:31-44   per ROW: FP, FN -> S=min, D=max(0,FN-FP), I=max(0,FP-FN)
  f1_overall_1sec       metrics.py:46-56   block max over ceil(N/block) blocks, then framewise F1
  er_overall_1sec       metrics.py:58-68   block max over floor(N/block) blocks (asymmetric!), then ER
  compute_scores        metrics.py:70-74

Everything before the last line of each function is integer counting; the restatement returns
the counts as well so the CUDA path can be compared count by count.  Pinned bit-exact against the
reference file itself by tests/golden/metrics_kat.npz.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(float).eps          # utils.py:4


def _as2d(a):
    a = np.asarray(a)
    if a.dtype == bool:
        a = a.astype(np.uint8)
    if a.ndim == 3:                  # utils.reshape_3Dto2D
        a = a.reshape(a.shape[0] * a.shape[1], a.shape[2])
    return a


def frame_counts(O, T):
    """-> dict of integer counts: TP, Nsys, Nref, S, D, I."""
    O, T = _as2d(O), _as2d(T)
    o1, t1 = (O == 1), (T == 1)
    tp = int(np.logical_and(o1, t1).sum())
    fp = np.logical_and(T == 0, o1).sum(1).astype(np.int64)
    fn = np.logical_and(t1, O == 0).sum(1).astype(np.int64)
    return dict(TP=tp, Nsys=int(O.sum()), Nref=int(T.sum()),
                S=int(np.minimum(fp, fn).sum()),
                D=int(np.maximum(0, fn - fp).sum()),
                I=int(np.maximum(0, fp - fn).sum()))


def _block_max(A, block, n_blocks):
    out = np.zeros((n_blocks, A.shape[1]))
    for i in range(n_blocks):
        out[i] = A[i * block:(i + 1) * block].max(axis=0)
    return out


def f1_from_counts(c):
    prec = float(c["TP"]) / float(c["Nsys"] + EPS)
    recall = float(c["TP"]) / float(c["Nref"] + EPS)
    return 2 * prec * recall / (prec + recall + EPS)


def er_from_counts(c):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.int64(c["S"] + c["D"] + c["I"]) / (np.float64(c["Nref"]) + 0.0)


def f1_overall_framewise(O, T):
    return f1_from_counts(frame_counts(O, T))


def er_overall_framewise(O, T):
    return er_from_counts(frame_counts(O, T))


def f1_overall_1sec(O, T, block_size):
    O, T = _as2d(O), _as2d(T)
    n = int(np.ceil(O.shape[0] / block_size))
    return f1_overall_framewise(_block_max(O, block_size, n), _block_max(T, block_size, n))


def er_overall_1sec(O, T, block_size):
    O, T = _as2d(O), _as2d(T)
    n = int(O.shape[0] / block_size)
    return er_overall_framewise(_block_max(O, block_size, n), _block_max(T, block_size, n))


def compute_scores(pred, y, frames_in_1_sec=50):
    return {"f1_overall_1sec": f1_overall_1sec(pred, y, frames_in_1_sec),
            "er_overall_1sec": er_overall_1sec(pred, y, frames_in_1_sec)}


def counts13(O, T, block):
    """The 13 integers `sedb200_threshold_counts` produces (include/sedb200.h), from the restatement above:
    frame-level TP, Nsys, Nref, S, D, I; block-max (ceil blocks, metrics.py:50) TP, Nsys, Nref; block-max (floor
    blocks, metrics.py:62) S, D, I, Nref."""
    O, T = _as2d(O), _as2d(T)
    fr = frame_counts(O, T)
    nc, nf = int(np.ceil(O.shape[0] / block)), int(O.shape[0] / block)
    f1 = frame_counts(_block_max(O, block, nc), _block_max(T, block, nc)) if nc else dict(TP=0, Nsys=0, Nref=0)
    er = frame_counts(_block_max(O, block, nf), _block_max(T, block, nf)) if nf else dict(S=0, D=0, I=0, Nref=0)
    return np.array([fr["TP"], fr["Nsys"], fr["Nref"], fr["S"], fr["D"], fr["I"], f1["TP"], f1["Nsys"], f1["Nref"],
                     er["S"], er["D"], er["I"], er["Nref"]], dtype=np.int64)
