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


def mbe_device(pcm: torch.Tensor, sr: int = SR, pad_mode: str = DEFAULT_PAD_MODE,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """Device-resident form.  pcm: CUDA float32 [S], [n_ch, S] or [n_clips, n_ch, S] ->
    [frames, 40], [frames, n_ch*40] or [n_clips, frames, n_ch*40] on the same device, enqueued on
    the current stream (no synchronisation)."""
    if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda):
        raise TypeError("mbe_device needs a CUDA tensor (no CPU fallback)")
    if pcm.dtype != torch.float32:
        raise TypeError("pcm must be float32 (feature.py:50 decodes to f32le)")
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
        _lib.check(_lib.lib().sedb200_logmel_f32(
            pcm.data_ptr(), n_clips, n_ch, S, int(sr), _lib.PAD_MODES[pad_mode], out.data_ptr(),
            _lib.current_stream_ptr()))
    return out


def mbe_batch(pcm: np.ndarray, sr: int = SR, pad_mode: str = DEFAULT_PAD_MODE,
              device: str | torch.device = "cuda") -> np.ndarray:
    """Host form for a batch: float32 [n_clips, n_ch, S] -> float32 [n_clips, frames, n_ch*40]."""
    x = np.ascontiguousarray(pcm, dtype=np.float32)
    if x.ndim != 3:
        raise ValueError("mbe_batch takes [n_clips, n_ch, S]")
    d = torch.from_numpy(x).to(device, non_blocking=False)
    return mbe_device(d, sr, pad_mode).cpu().numpy()


def _mbe(y: np.ndarray, sr: int = SR) -> np.ndarray:
    """feature._mbe (feature.py:55-59): mono float32 PCM -> log-mel (frames, 40), host in / host out."""
    y = np.ascontiguousarray(y, dtype=np.float32)
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
