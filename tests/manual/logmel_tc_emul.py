"""CPU emulation of the data flow of logmel_tc_kernel (csrc/logmel_tc.cu): the 2048-point real DFT of a Hann-windowed
frame as TWO batched GEMMs on fp16 hi/lo operand planes with fp32 accumulation,

    n = 32 n1 + n2,  k = k1 + 64 k2
    stage 1   rows (frame, n2), K = n1 (64 reals)      -> Y[n2][k1], k1 = 0..32 (Hermitian-packed: 64 reals)
    twiddle   Z[n2][k1] = Y[n2][k1] W_2048^(n2 k1)      (fp32, CUDA cores)
    stage 2   rows (frame, k1 = 0..31), K = (n2, re/im) -> X[k1 + 64 k2], k2 = 0..31   (k2 >= 16: mirror bins)
    special   rows (frame), K = n2 (Y[n2][32], real)    -> X[32 + 64 k2], k2 = 0..15

and the same |X|^2 -> mel -> log tail.  Used to settle operand precision (how many split terms) and the index maps
before any CUDA was written; run by hand:  python tests/manual/logmel_tc_emul.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import logmel_ref as L  # noqa: E402

N1, N2 = 64, 32
two_pi = 2 * np.pi


def split16(x):
    hi = x.astype(np.float16)
    lo = (x.astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def mm3(a, b, terms=3):
    """a [M,K] fp32 data, b [K,N] float64 constants -> fp32 [M,N] through fp16 hi/lo planes."""
    ah, al = split16(a)
    bh, bl = split16(b.astype(np.float32))
    out = ah @ bh
    if terms >= 2:
        out = out + al @ bh
    if terms >= 3:
        out = out + ah @ bl
    return out.astype(np.float32)


def matrices():
    n1 = np.arange(N1)[:, None]
    c = np.arange(32)[None, :]
    b1 = np.zeros((N1, 64))
    b1[:, :32] = np.cos(two_pi * n1 * c / 64)
    b1[:, 32:] = -np.sin(two_pi * n1 * c / 64)
    b1[:, 32] = np.cos(np.pi * n1[:, 0])                      # slot of Im Y[0] (= 0) carries Y[32]
    n2 = np.arange(N2)[:, None]
    th = two_pi * n2 * c / 32
    b2 = np.zeros((64, 64))
    b2[0::2, :32] = np.cos(th)
    b2[1::2, :32] = np.sin(th)
    b2[0::2, 32:] = -np.sin(th)
    b2[1::2, 32:] = np.cos(th)
    k2 = np.arange(16)[None, :]
    ph = two_pi * n2 * (2 * k2 + 1) / 64
    b2s = np.concatenate([np.cos(ph), -np.sin(ph)], axis=1)  # [32, 32]
    return b1, b2, b2s


def frames_of(y, pad_mode):
    yp = np.pad(np.asarray(y, np.float32), L.NFFT // 2, mode=pad_mode)
    nfr = 1 + (yp.shape[0] - L.NFFT) // L.HOP
    idx = np.arange(L.NFFT)[None, :] + L.HOP * np.arange(nfr)[:, None]
    return yp[idx]                                              # [frames, 2048]


def mbe_tc(y, pad_mode="constant", terms=3, scale=True):
    fr = frames_of(y, pad_mode)
    F = fr.shape[0]
    w = (0.5 - 0.5 * np.cos(two_pi * np.arange(L.NFFT) / L.NFFT)).astype(np.float32)
    yw = (fr * w).astype(np.float32)
    if scale:                                                   # per-frame power of two: max |yw| in [256, 512)
        mx = np.abs(yw).max(axis=1)
        e = np.where(mx > 0, 8 - np.floor(np.log2(np.maximum(mx, 1e-45))), 0.0)
        sc = np.exp2(e).astype(np.float32)
    else:
        sc = np.ones(F, np.float32)
    yw = yw * sc[:, None]
    b1, b2, b2s = matrices()
    a1 = yw.reshape(F, N1, N2).transpose(0, 2, 1).reshape(F * N2, N1)      # rows (f, n2), K = n1
    d1 = mm3(a1, b1, terms).reshape(F, N2, 64)
    yre, yim = d1[:, :, :32].copy(), d1[:, :, 32:].copy()
    y32 = yim[:, :, 0].copy()
    yim[:, :, 0] = 0
    n2 = np.arange(N2)[:, None]
    k1 = np.arange(32)[None, :]
    ang = two_pi * n2 * k1 / 2048
    tc, ts = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    zre = (yre * tc + yim * ts).astype(np.float32)             # (re + i im)(c - i s)
    zim = (yim * tc - yre * ts).astype(np.float32)
    a2 = np.empty((F, 32, 64), np.float32)                     # rows (f, k1), K = (n2, re/im)
    a2[:, :, 0::2] = zre.transpose(0, 2, 1)
    a2[:, :, 1::2] = zim.transpose(0, 2, 1)
    d2 = mm3(a2.reshape(F * 32, 64), b2, terms).reshape(F, 32, 64)
    pw = (d2[:, :, :32] ** 2 + d2[:, :, 32:] ** 2).astype(np.float32)      # [f, k1, k2]
    d2s = mm3(y32, b2s, terms)                                 # [F, 32]
    pws = (d2s[:, :16] ** 2 + d2s[:, 16:] ** 2).astype(np.float32)
    P = np.zeros((F, 1025), np.float32)
    for kk1 in range(32):
        for kk2 in range(32):
            k = kk1 + 64 * kk2
            if kk2 >= 16:
                k = 2048 - k
            if kk1 == 0 and kk2 > 16:
                continue
            P[:, k] = pw[:, kk1, kk2]
    for kk2 in range(16):
        P[:, 32 + 64 * kk2] = pws[:, kk2]
    mel = (P @ L.mel_filterbank().T.astype(np.float32)).astype(np.float32)
    with np.errstate(divide="ignore"):
        out = np.log(mel) - (2 * np.log(2.0) * np.log2(sc))[:, None]
    return out.astype(np.float32), P / (sc[:, None] ** 2)


def close(got, want, floor=25.0):
    err = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1.0)
    audible = want >= want.max(axis=-1, keepdims=True) - floor
    return float(np.where(audible, err, 0.0).max())


if __name__ == "__main__":
    # index map first: against the float64 rFFT with all split terms the power spectrum must agree to ~1e-6
    y = L.synth_clip(3, 5000, 1, "mix")[0]
    _, P = mbe_tc(y)
    ref = np.abs(L.stft(y).T.astype(np.complex128)) ** 2
    print("power spectrum rel err (3 terms):", float(np.abs(P - ref).max() / ref.max()))
    for terms in (3, 2, 1):
        worst = 0.0
        for kind in ("mix", "noise", "chirp"):
            for n in (1, 1000, 1025, 4096, 44100):
                for pm in ("constant", "reflect"):
                    y = L.synth_clip(n % 97, n, 1, kind)[0]
                    got, _ = mbe_tc(y, pm, terms)
                    worst = max(worst, close(got, L.mbe(y, pad_mode=pm)))
        print(f"terms={terms}: worst gate value {worst:.3e} (gate 1e-4)")
    # quiet input and loud input: the per-frame scale keeps fp16 in range
    for amp in (1e-4, 1.0, 3e4):
        y = (L.synth_clip(5, 8000, 1, "mix")[0] * amp).astype(np.float32)
        for sc in (True, False):
            with np.errstate(over="ignore", invalid="ignore"):
                got, _ = mbe_tc(y, "constant", 3, sc)
            print(f"amp={amp:g} scale={sc}: {close(got, L.mbe(y)):.3e}")
