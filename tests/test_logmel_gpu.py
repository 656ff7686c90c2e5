"""GPU parity of the log-mel kernel (through the C ABI) against the CPU oracle.
Gate (BASELINE.json north_star): log-mel within 1e-4 relative in fp32.  The comparison is
|got - want| <= 1e-4 * max(|want|, 1): relative where the log value is O(1) or larger, absolute 1e-4
(= 1e-4 relative on the mel ENERGY) where log-mel passes through zero.  Bands more than 25 nats
(-108 dB) below the loudest band of their frame are skipped: that is beyond what fp32 frame arithmetic
can resolve (the reference gets there only because librosa's FFT runs in float64) and only occurs for
noise-free synthetic inputs such as a 1-sample clip under reflect padding."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_ref as L

pytestmark = pytest.mark.gpu
RTOL = 1e-4
FLOOR_NATS = 25.0


def close(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    err = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1.0)
    audible = want >= want.max(axis=-1, keepdims=True) - FLOOR_NATS
    assert audible.mean() > 0.02
    return float(np.where(audible, err, 0.0).max())


@pytest.fixture(scope="module")
def feat(built_lib):
    assert torch.cuda.is_available()
    from sed_crnn_b200 import feature
    return feature


@pytest.mark.parametrize("kind", ["mix", "noise", "chirp"])
@pytest.mark.parametrize("n", [1, 1000, 1023, 1024, 1025, 2048, 2049, 4096, 44100, 2 * 44100 + 1, 1024 * 37])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_mbe_parity_mono(feat, kind, n, pad_mode):
    y = L.synth_clip(n % 97, n, 1, kind)[0]
    want = L.mbe(y, pad_mode=pad_mode)
    got = feat.mbe_device(torch.from_numpy(y).cuda(), pad_mode=pad_mode).cpu().numpy()
    assert close(got, want) <= RTOL


def test_drop_in_signature(feat):
    y = L.synth_clip(11, 3 * 44100, 1, "mix")[0]
    got = feat._mbe(y, feat.SR)
    assert isinstance(got, np.ndarray) and got.shape == (1 + len(y) // 1024, 40)
    assert close(got, L.mbe(y)) <= RTOL
    assert (feat.SR, feat.NFFT, feat.HOP, feat.NB_MEL) == (44100, 2048, 1024, 40)


@pytest.mark.parametrize("n_ch,n", [(2, 30001), (2, 30000), (3, 5000)])
def test_multichannel_batch_layout(feat, n_ch, n):
    clips = np.stack([L.synth_clip(20 + i, n, n_ch, "mix") for i in range(3)])
    got = feat.mbe_device(torch.from_numpy(clips).cuda()).cpu().numpy()
    assert got.shape == (3, 1 + n // 1024, n_ch * 40)
    for i in range(3):
        assert close(got[i], L.mbe_multichannel(clips[i])) <= RTOL


def test_golden_fixture(feat, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_oracle.npz"))
    for name in ("mix_1s", "noise_odd", "chirp_stereo", "short"):
        for pm in ("constant", "reflect"):
            got = feat.mbe_device(torch.from_numpy(g[name + "_pcm"]).cuda(), pad_mode=pm).cpu().numpy()
            assert close(got, g[f"{name}_{pm}"]) <= RTOL, (name, pm)


def test_other_sample_rate(feat):
    y = L.synth_clip(5, 22050, 1, "noise")[0]
    got = feat.mbe_device(torch.from_numpy(y).cuda(), sr=22050).cpu().numpy()
    assert close(got, L.mbe(y, sr=22050)) <= RTOL


def test_silence_and_errors(feat):
    z = feat.mbe_device(torch.zeros(5000, device="cuda")).cpu().numpy()
    assert np.all(np.isneginf(z))
    with pytest.raises(ValueError):
        feat._mbe(np.zeros(0, np.float32), 44100)
    with pytest.raises(TypeError):
        feat.mbe_device(torch.zeros(10))


def test_full_size_properties(feat):
    """BASELINE size (3-min stereo clip): size-independent checks -- frame-shift invariance (the same
    audio delayed by one hop gives the same interior frames) and agreement with the oracle on a
    strided sample of frames."""
    n = 180 * 44100
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal((2, n + 1024))).astype(np.float32)
    d = torch.from_numpy(y).cuda()
    a = feat.mbe_device(d[:, 1024:].contiguous())
    b = feat.mbe_device(d[:, :n].contiguous())
    assert a.shape == (7752, 80)
    assert torch.equal(a[1:-2], b[2:-1])
    idx = np.arange(0, 7752, 517)
    want = L.mbe(y[1, 1024:])[idx]
    assert close(a[:, 40:].cpu().numpy()[idx], want) <= RTOL


def test_determinism(feat):
    y = torch.from_numpy(L.synth_clip(1, 400000, 2, "mix")).cuda()
    a = feat.mbe_device(y)
    for _ in range(3):
        assert torch.equal(a, feat.mbe_device(y))


def test_standard_scaler_matches_sklearn(feat):
    """feature.py:127-129: StandardScaler fit on train frames, applied to train and test."""
    sk = pytest.importorskip("sklearn.preprocessing")
    rng = np.random.default_rng(3)
    Xtr = (rng.standard_normal((20011, 40)) * rng.uniform(0.5, 3, 40) + rng.uniform(-8, 2, 40)).astype(np.float32)
    Xtr[:, 7] = 1.25                                     # constant column -> scale 1
    Xte = (rng.standard_normal((999, 40)) * 2 - 3).astype(np.float32)
    ref = sk.StandardScaler()
    want_tr, want_te = ref.fit_transform(Xtr.copy()), ref.transform(Xte.copy())
    mine = feat.StandardScaler()
    got_tr, got_te = mine.fit_transform(Xtr), mine.transform(Xte)
    np.testing.assert_allclose(mine.mean_, ref.mean_, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mine.var_, ref.var_, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mine.scale_, ref.scale_, rtol=1e-5)
    assert mine.n_samples_seen_ == 20011 and got_tr.dtype == np.float32
    np.testing.assert_allclose(got_tr, want_tr, rtol=0, atol=2e-5)
    np.testing.assert_allclose(got_te, want_te, rtol=0, atol=2e-5)
    d = mine.transform(torch.from_numpy(Xte).cuda())
    assert d.is_cuda and torch.equal(d.cpu(), torch.from_numpy(got_te))


@pytest.mark.parametrize("n_ch,n", [(1, 1), (1, 2047), (2, 30001), (1, 44100), (2, 1024 * 9)])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_int16_ingest_is_bit_identical_to_float(feat, n_ch, n, pad_mode):
    """sedb200_logmel_i16: a 16-bit sample s stands for the float32 value s / 32768, which is exact,
    so the int16 path must give the very bits of the float32 path on the converted signal -- including odd clip
    lengths (unaligned frame starts of the second channel) and both padding rules."""
    rng = np.random.default_rng(7 + n)
    pcm16 = rng.integers(-32768, 32768, size=(n_ch, n), dtype=np.int16)
    pcm16[pcm16 == 0] = 1                                   # digital silence is log(0) = -inf (feature.py:59)
    as_float = (pcm16.astype(np.float32) / np.float32(32768.0))
    a = feat.mbe_device(torch.from_numpy(pcm16).cuda(), pad_mode=pad_mode)
    b = feat.mbe_device(torch.from_numpy(as_float).cuda(), pad_mode=pad_mode)
    assert torch.equal(a, b)
    want = L.mbe_multichannel(as_float, pad_mode=pad_mode)
    assert close(a.cpu().numpy(), want) <= RTOL


def test_int16_host_drop_in(feat):
    y = L.synth_clip(5, 2 * 44100, 1, "mix")[0]
    y16 = np.round(np.clip(y, -1, 1) * 32767).astype(np.int16)
    got = feat._mbe(y16, feat.SR)
    want = L.mbe(y16.astype(np.float32) / np.float32(32768.0))
    assert got.shape == want.shape and close(got, want) <= RTOL


def test_against_third_party_golden(feat, golden_dir):
    """The CUDA kernel against vectors of `transformers.audio_utils` (oracle/make_golden_logmel_thirdparty.py): an
    implementation that shares no code with this repository or its oracle."""
    g = np.load(os.path.join(golden_dir, "logmel_thirdparty.npz"))
    for name in ("mix_1s", "noise_odd", "chirp_7k", "short"):
        for pm in ("constant", "reflect"):
            got = feat.mbe_device(torch.from_numpy(g[name + "_pcm"]).cuda(), pad_mode=pm).cpu().numpy()
            assert close(got, g[f"{name}_{pm}"].astype(np.float64)) <= RTOL, (name, pm)
