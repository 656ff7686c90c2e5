"""Oracle self-checks for the log-mel restatement (CPU only).  The reference holds no golden vector for
this leg (librosa is an absent third-party dependency) -- these are cross-checks, not pins."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_ref as L


def test_mel_filterbank_known_answers():
    m = L.mel_filterbank()
    assert m.shape == (40, 1025) and m.dtype == np.float32
    assert int((m != 0).sum()) == 1945
    assert int((m != 0).sum(0).max()) == 2
    np.testing.assert_allclose(m.sum(), 1.857314, rtol=1e-6)
    np.testing.assert_allclose(m.max(), 0.0101148, rtol=1e-5)
    e = L.mel_band_edges()
    np.testing.assert_allclose(e[:4], [0.0, 97.549, 195.097, 292.646], atol=1e-3)
    np.testing.assert_allclose(e[-3:], [18031.37, 19939.70, 22050.0], atol=1e-2)


def test_mel_filterbank_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(1025, 0.0, 22050.0, 40, 44100, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(ref - L.mel_filterbank()).max() < 2e-8


@pytest.mark.parametrize("n", [1000, 2048, 44100, 2 * 44100 + 1, 1024 * 7])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_mbe_vs_torch_stft(n, pad_mode):
    y = L.synth_clip(3, n, 1, "mix")[0]
    o = L.mbe(y, pad_mode=pad_mode)
    assert o.shape == (1 + n // 1024, 40) and o.dtype == np.float32
    if pad_mode == "reflect" and n <= 1024:
        return                                           # torch.stft refuses reflect pad >= length
    w = torch.hann_window(2048, periodic=True, dtype=torch.float64)
    S = torch.stft(torch.from_numpy(y).double(), 2048, 1024, window=w, center=True, pad_mode=pad_mode,
                   return_complex=True)
    P = (S.abs() ** 2).numpy().astype(np.float32)
    o2 = np.log(L.mel_filterbank() @ P).T
    np.testing.assert_allclose(o, o2, rtol=2e-5, atol=2e-5)


def test_pad_modes_differ_only_at_the_ends():
    y = L.synth_clip(4, 20000, 1, "noise")[0]
    a, b = L.mbe(y, pad_mode="constant"), L.mbe(y, pad_mode="reflect")
    assert np.array_equal(a[1:-1], b[1:-1]) and not np.array_equal(a[0], b[0])


def test_multichannel_layout_matches_split_multi_channels():
    y = L.synth_clip(5, 5000, 2, "mix")
    o = L.mbe_multichannel(y)
    assert o.shape == (5, 80)
    np.testing.assert_array_equal(o[:, :40], L.mbe(y[0]))
    np.testing.assert_array_equal(o[:, 40:], L.mbe(y[1]))


def test_empty_raises_and_silence_is_minus_inf():
    with pytest.raises(ValueError):
        L.mbe(np.zeros(0, np.float32))
    assert np.all(np.isneginf(L.mbe(np.zeros(3000, np.float32))))


def test_frozen_anchor(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_oracle.npz"))
    for name in ("mix_1s", "noise_odd", "chirp_stereo", "short"):
        for pm in ("constant", "reflect"):
            got = L.mbe_multichannel(g[name + "_pcm"], pad_mode=pm)
            np.testing.assert_allclose(got, g[f"{name}_{pm}"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(g["mel_fb"], L.mel_filterbank())


def test_third_party_golden(golden_dir):
    """Vectors produced by `transformers.audio_utils` (an independent librosa-compatible implementation; generating
    script: oracle/make_golden_logmel_thirdparty.py).  Not the reference's own numbers -- librosa is absent -- but a
    pin that does not come from this repository's code."""
    g = np.load(os.path.join(golden_dir, "logmel_thirdparty.npz"))
    for name in ("mix_1s", "noise_odd", "chirp_7k", "short"):
        for pm in ("constant", "reflect"):
            want = g[f"{name}_{pm}"]
            got = L.mbe(g[name + "_pcm"], pad_mode=pm)
            assert got.shape == want.shape
            err = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1.0)
            assert err.max() <= 5e-6, (name, pm, err.max())
