"""CPU restatement of the reference log-mel front end.  TEST INFRASTRUCTURE (see oracle/__init__).

Follows /root/reference/feature.py:55-59 (`_mbe`) line by line:

    s          = librosa.stft(y, n_fft=2048, hop_length=1024)      feature.py:56
    power_spec = np.abs(s) ** 2                                    feature.py:57
    mel_b      = librosa.filters.mel(sr=sr, n_fft=2048, n_mels=40) feature.py:58
    return np.log(np.dot(mel_b, power_spec)).T                     feature.py:59

librosa itself is a third-party dependency that is NOT in /root/reference (requirements.txt:4
pins `librosa==0.7.0`, but feature.py:58 uses the keyword-only `filters.mel(sr=...)` form that
only exists in librosa >= 0.10, so the pin is stale; effectively unpinned).  What is restated
here is librosa's published algorithm:

  * stft(center=True): pad n_fft//2 samples on both sides (`pad_mode` 'constant' in
    librosa >= 0.10, 'reflect' before), frames of n_fft at hop, multiplied by the periodic Hann
    window `scipy.signal.get_window('hann', n_fft, fftbins=True)` in float64, rFFT in float64,
    result stored as complex64.
  * filters.mel(htk=False, norm='slaney', fmin=0, fmax=sr/2, dtype=float32): Slaney mel scale,
    triangular filters, area normalisation 2/(f[i+2]-f[i]).

PARITY UNPINNED by the reference (it holds no tests / golden vectors).  Cross-checks live in
tests/test_oracle_logmel.py (torchaudio.functional.melscale_fbanks, torch.stft), together with committed vectors of an
independent third-party implementation (transformers.audio_utils; oracle/make_golden_logmel_thirdparty.py ->
tests/golden/logmel_thirdparty.npz), which this restatement matches to 5e-7.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

SR = 44_100          # feature.py:29
NFFT = 2048          # feature.py:30
HOP = NFFT // 2      # feature.py:31
NB_MEL = 40          # feature.py:32


# --------------------------------------------------------------------------- mel filterbank
def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    big = f >= min_log_hz
    out = np.array(mels, dtype=np.float64, copy=True)
    out[big] = min_log_mel + np.log(f[big] / min_log_hz) / logstep
    return out


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    big = m >= min_log_mel
    out = np.array(freqs, dtype=np.float64, copy=True)
    out[big] = min_log_hz * np.exp(logstep * (m[big] - min_log_mel))
    return out


def mel_band_edges(sr: int = SR, n_mels: int = NB_MEL) -> np.ndarray:
    """n_mels+2 band-edge frequencies in Hz (float64)."""
    lo = _hz_to_mel_slaney(np.array([0.0]))[0]
    hi = _hz_to_mel_slaney(np.array([sr / 2.0]))[0]
    return _mel_to_hz_slaney(np.linspace(lo, hi, n_mels + 2))


def mel_filterbank(sr: int = SR, n_fft: int = NFFT, n_mels: int = NB_MEL) -> np.ndarray:
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels) -> float32 [n_mels, 1+n_fft//2]."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_band_edges(sr, n_mels)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))      # float64 -> float32 store
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]                                # in-place, rounds to float32
    return weights


# --------------------------------------------------------------------------- STFT
def n_frames(n_samples: int, hop: int = HOP) -> int:
    return 1 + n_samples // hop


def stft(y: np.ndarray, n_fft: int = NFFT, hop: int = HOP, pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(y, n_fft, hop_length=hop) -> complex64 [1+n_fft//2, 1+len(y)//hop]."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("oracle stft takes a 1-D signal")
    if y.shape[0] == 0:
        raise ValueError("empty signal")
    if pad_mode not in ("constant", "reflect"):
        raise ValueError(f"pad_mode {pad_mode!r}")
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)     # float64, periodic
    nfr = 1 + (yp.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(nfr)[None, :]
    frames = yp[idx]                                               # [n_fft, frames], y dtype
    spec = scipy.fft.rfft(win[:, None] * frames, axis=0)           # float64 arithmetic
    return spec.astype(np.complex64)


def mbe(y: np.ndarray, sr: int = SR, pad_mode: str = "constant") -> np.ndarray:
    """feature._mbe restated: float32 [frames, 40] natural-log mel-band energies."""
    s = stft(np.asarray(y, dtype=np.float32), NFFT, HOP, pad_mode)
    power_spec = np.abs(s) ** 2                                    # float32
    mel_b = mel_filterbank(sr, NFFT, NB_MEL)                       # float32
    with np.errstate(divide="ignore"):
        return np.log(np.dot(mel_b, power_spec)).T


def mbe_multichannel(y: np.ndarray, sr: int = SR, pad_mode: str = "constant") -> np.ndarray:
    """[n_ch, S] -> [frames, n_ch*40]; channel-major on the feature axis, the layout that
    /root/reference/utils.py:15-25 (`split_multi_channels`) undoes."""
    y = np.asarray(y, dtype=np.float32)
    if y.ndim == 1:
        return mbe(y, sr, pad_mode)
    return np.concatenate([mbe(c, sr, pad_mode) for c in y], axis=1)


# --------------------------------------------------------------------------- synthetic audio
def synth_clip(clip_id: int, n_samples: int, n_ch: int = 1, kind: str = "mix") -> np.ndarray:
    """Seeded synthetic PCM in [-1,1]: 0.1*N(0,1) noise floor plus three
    sinusoids (kind='mix'), noise only ('noise'), or a linear chirp over the floor ('chirp')."""
    rng = np.random.default_rng(1000 + clip_id)
    t = np.arange(n_samples, dtype=np.float64) / SR
    out = np.empty((n_ch, n_samples), dtype=np.float32)
    for c in range(n_ch):
        x = 0.1 * rng.standard_normal(n_samples)
        if kind == "mix":
            ph = rng.uniform(0, 2 * np.pi, 3)
            x += 0.3 * np.sin(2 * np.pi * 440.0 * t + ph[0])
            x += 0.1 * np.sin(2 * np.pi * 3000.0 * t + ph[1])
            x += 0.03 * np.sin(2 * np.pi * 12000.0 * t + ph[2])
        elif kind == "chirp":
            dur = max(n_samples / SR, 1e-3)
            k = (20000.0 - 50.0) / dur
            x += 0.3 * np.sin(2 * np.pi * (50.0 * t + 0.5 * k * t * t))
        elif kind != "noise":
            raise ValueError(kind)
        out[c] = np.clip(x, -1.0, 1.0).astype(np.float32)
    return out
