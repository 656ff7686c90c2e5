"""Drop-in for the log-mel front end of /root/reference/feature.py.

Same names and call surface as the reference module (feature.py:29-32, 55-59):

    SR, NFFT, HOP, NB_MEL
    _mbe(y: float32[S], sr) -> float32[1 + S//1024, 40]

plus the batched / device-resident forms the B200 path needs (`mbe_batch`, `mbe_device`).  All of
them run the sm_100a kernel in libsedb200.so; there is no CPU implementation here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

SR = 44_100            # feature.py:29
NFFT = 2048            # feature.py:30
HOP = NFFT // 2        # feature.py:31
NB_MEL = 40            # feature.py:32

#: librosa >= 0.10 (the only line that accepts feature.py:58's keyword call) pads with zeros;
#: set to "reflect" to reproduce librosa < 0.10.
DEFAULT_PAD_MODE = "constant"


def n_frames(n_samples: int) -> int:
    return 1 + n_samples // HOP


LOGMEL_KERNELS = {"auto": 0, "fp32": 1, "tc": 2}      # include/sedb200.h: SEDB200_LOGMEL_*


def mbe_device(pcm: torch.Tensor, sr: int = SR, pad_mode: str = DEFAULT_PAD_MODE,
               out: torch.Tensor | None = None, kernel: str = "auto") -> torch.Tensor:
    """Device-resident form.  pcm: CUDA float32 (or int16, value = s / 32768) [S], [n_ch, S] or [n_clips, n_ch, S] ->
    [frames, 40], [frames, n_ch*40] or [n_clips, frames, n_ch*40] on the same device, enqueued on
    the current stream (no synchronisation).  `kernel`: "fp32" (CUDA-core FFT), "tc" (tcgen05 DFT-as-GEMM) or "auto"
    (the library default / SEDB200_LOGMEL_KERNEL)."""
    if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda):
        raise TypeError("mbe_device needs a CUDA tensor (no CPU fallback)")
    if pcm.dtype not in (torch.float32, torch.int16):
        raise TypeError("pcm must be float32 (feature.py:50 decodes to f32le) or int16 (s16le / WAV samples, "
                        "value = s / 32768)")
    shape = pcm.shape
    if pcm.dim() == 1:
        n_clips, n_ch, S = 1, 1, shape[0]
    elif pcm.dim() == 2:
        n_clips, n_ch, S = 1, shape[0], shape[1]
    elif pcm.dim() == 3:
        n_clips, n_ch, S = shape
    else:
        raise ValueError("pcm must be [S], [n_ch,S] or [n_clips,n_ch,S]")
    if S < 1:
        raise ValueError("empty signal")
    pcm = pcm.contiguous()
    nfr = n_frames(S)
    oshape = (nfr, n_ch * NB_MEL) if pcm.dim() < 3 else (n_clips, nfr, n_ch * NB_MEL)
    if out is None:
        out = torch.empty(oshape, dtype=torch.float32, device=pcm.device)
    elif tuple(out.shape) != oshape or out.dtype != torch.float32 or not out.is_contiguous() \
            or out.device != pcm.device:
        raise ValueError(f"out must be contiguous float32 {oshape} on {pcm.device}")
    with torch.cuda.device(pcm.device):
        fn = _lib.lib().sedb200_logmel_i16_k if pcm.dtype == torch.int16 else _lib.lib().sedb200_logmel_f32_k
        _lib.check(fn(pcm.data_ptr(), n_clips, n_ch, S, int(sr), _lib.PAD_MODES[pad_mode], out.data_ptr(),
                      _lib.current_stream_ptr(), LOGMEL_KERNELS[kernel]))
    return out


def mbe_batch(pcm: np.ndarray, sr: int = SR, pad_mode: str = DEFAULT_PAD_MODE,
              device: str | torch.device = "cuda") -> np.ndarray:
    """Host form for a batch: float32 (or int16) [n_clips, n_ch, S] -> float32 [n_clips, frames, n_ch*40]."""
    pcm = np.asarray(pcm)
    x = np.ascontiguousarray(pcm, dtype=np.int16 if pcm.dtype == np.int16 else np.float32)
    if x.ndim != 3:
        raise ValueError("mbe_batch takes [n_clips, n_ch, S]")
    d = torch.from_numpy(x).to(device, non_blocking=False)
    return mbe_device(d, sr, pad_mode).cpu().numpy()


def _mbe(y: np.ndarray, sr: int = SR) -> np.ndarray:
    """feature._mbe (feature.py:55-59): mono float32 PCM -> log-mel (frames, 40), host in / host out."""
    y = np.asarray(y)
    y = np.ascontiguousarray(y, dtype=np.int16 if y.dtype == np.int16 else np.float32)   # int16: s16le samples
    if y.ndim != 1:
        raise ValueError("_mbe takes a mono 1-D signal (feature.py:45 decodes with -ac 1)")
    if y.shape[0] == 0:
        raise ValueError("empty signal")
    d = torch.from_numpy(y).cuda()
    return mbe_device(d, sr, DEFAULT_PAD_MODE).cpu().numpy()


def mel_filterbank(sr: int = SR) -> np.ndarray:
    """The float32 [40, 1025] filterbank the kernel uses (host-side table, for inspection)."""
    out = np.empty((NB_MEL, NFFT // 2 + 1), dtype=np.float32)
    _lib.check(_lib.lib().sedb200_mel_filterbank(int(sr), out.ctypes.data))
    return out


def rasterize_labels(starts_s, ends_s, n_frames: int, sr: int = SR, hop: int = HOP,
                     out: torch.Tensor | None = None, col: int = 0) -> torch.Tensor:
    """feature.py:88-93 on the device: frame labels [n_frames, 1] (CUDA float32) from event start / end times in
    seconds, `lbl[floor(start*SR/HOP) : ceil(end*SR/HOP), 0] = 1.0` with Python slice clamping."""
    a = torch.as_tensor(np.asarray(starts_s, dtype=np.float64)).reshape(-1)
    b = torch.as_tensor(np.asarray(ends_s, dtype=np.float64)).reshape(-1)
    if a.numel() != b.numel():
        raise ValueError("starts and ends differ in length")
    if out is None:
        out = torch.zeros(n_frames, 1, dtype=torch.float32, device="cuda")
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == n_frames):
        raise ValueError("out must be a contiguous CUDA float32 [n_frames, n_lab] tensor")
    ev = torch.stack([a, b]).to(out.device).contiguous()
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().sedb200_rasterize_labels(ev[0].data_ptr(), ev[1].data_ptr(), a.numel(), int(sr), int(hop),
                                                       n_frames, out.shape[1], col, out.data_ptr(),
                                                       _lib.current_stream_ptr()))
    return out


class StandardScaler:
    """Drop-in for the `sklearn.preprocessing.StandardScaler` calls of feature.py:127-129 (fit on the training
    frames, apply to both splits), computed on the GPU: deterministic per-bin sum / sum of squares reduction,
    float64 statistics (`mean_`, `var_`, `scale_`, `n_samples_seen_`), float32 output.  Accepts CUDA tensors
    (returns CUDA tensors) or numpy arrays (returns numpy)."""

    def __init__(self):
        self.mean_ = self.var_ = self.scale_ = None
        self.n_samples_seen_ = 0
        self._dev = None

    @staticmethod
    def _to_dev(X):
        if isinstance(X, torch.Tensor):
            if not X.is_cuda:
                raise TypeError("StandardScaler needs CUDA tensors or numpy arrays (no CPU fallback)")
            return X.contiguous().float(), False
        return torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda(), True

    def fit(self, X):
        x, _ = self._to_dev(X)
        if x.dim() != 2 or x.shape[0] < 1:
            raise ValueError("expected a non-empty 2-D array [frames, features]")
        rows, cols = x.shape
        L = _lib.lib()
        stats = torch.empty(3, cols, dtype=torch.float64, device=x.device)
        nbytes = int(L.sedb200_standardize_scratch_bytes(rows, cols))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.sedb200_standardize_fit(x.data_ptr(), rows, cols, stats[0].data_ptr(), stats[1].data_ptr(),
                                                 stats[2].data_ptr(), scratch.data_ptr(), nbytes,
                                                 _lib.current_stream_ptr()))
        self._dev = stats
        host = stats.cpu().numpy()
        self.mean_, self.var_, self.scale_ = host[0], host[1], host[2]
        self.n_samples_seen_ = rows
        return self

    def transform(self, X):
        if self._dev is None:
            raise RuntimeError("StandardScaler.transform called before fit")
        x, was_numpy = self._to_dev(X)
        if x.dim() != 2 or x.shape[1] != self._dev.shape[1]:
            raise ValueError("feature count differs from the fitted data")
        stats = self._dev.to(x.device)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sedb200_standardize_apply(x.data_ptr(), x.shape[0], x.shape[1], stats[0].data_ptr(),
                                                            stats[2].data_ptr(), out.data_ptr(),
                                                            _lib.current_stream_ptr()))
        return out.cpu().numpy() if was_numpy else out

    def fit_transform(self, X):
        return self.fit(X).transform(X)


# ----------------------------------------------------------------------------------------------- fold packs
FOLD_PACK_FMT = "mbe_mon_fold{}.npz"         # feature.py:131; read back by decorte_datamodule._load_all_npz / sed.load_all_npz


def _dev2d(a) -> torch.Tensor:
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    if t.dim() == 1:
        t = t[:, None]
    return t.to("cuda", torch.float32, non_blocking=True)


def pack_folds(per_video: dict, cache_dir: str, n_folds: int | None = None) -> list[str]:
    """feature.py:109-133 -- the step between the two halves of the hot path, with the arithmetic on the device.

    per_video: {video_name: (mbe [frames, n_feat], lbl [frames, 1], fold_id)} in the reference's iteration order
    (numpy arrays or CUDA tensors, e.g. straight from `mbe_device` / `rasterize_labels`).  For every fold f:
    test = the videos of fold f concatenated in dict order, train = all the others; a `StandardScaler` is fitted on
    the training frames and applied to both (feature.py:127-129); the pack is written with `np.savez` positional
    arrays, i.e. keys `arr_0..arr_3` = X_train, Y_train, X_test, Y_test, float32, into
    `cache_dir/mbe_mon_fold{f+1}.npz` -- the file format `_load_all_npz` (decorte_datamodule.py:24-34) and
    `load_all_npz` (sed.py:115-125) read.  Returns the paths written."""
    import os
    if not per_video:
        raise ValueError("pack_folds: no videos")
    fold_k = (max(v[2] for v in per_video.values()) + 1) if n_folds is None else int(n_folds)     # feature.py:112
    dev = {k: (_dev2d(m), _dev2d(l), int(f)) for k, (m, l, f) in per_video.items()}
    paths = []
    for f in range(fold_k):
        tr = [(m, l) for (m, l, fold) in dev.values() if fold != f]
        te = [(m, l) for (m, l, fold) in dev.values() if fold == f]
        if not tr or not te:
            raise ValueError(f"pack_folds: fold {f} has no {'training' if not tr else 'test'} video")
        X_train, Y_train = torch.cat([m for m, _ in tr]), torch.cat([l for _, l in tr])
        X_test, Y_test = torch.cat([m for m, _ in te]), torch.cat([l for _, l in te])
        scaler = StandardScaler()
        X_train = scaler.fit_transform(X_train)
        X_test = scaler.transform(X_test)
        out_fold = os.path.join(cache_dir, FOLD_PACK_FMT.format(f + 1))
        np.savez(out_fold, X_train.cpu().numpy(), Y_train.cpu().numpy(), X_test.cpu().numpy(), Y_test.cpu().numpy())
        paths.append(out_fold)
    return paths
