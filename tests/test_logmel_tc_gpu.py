"""GPU parity of the tensor-core log-mel kernel (csrc/logmel_tc.cu, `kernel="tc"`: the 2048-point DFT as two batched
tcgen05 GEMMs on fp16 hi / lo planes) through the C ABI against the CPU oracle -- the same cases and the same gate as
tests/test_logmel_gpu.py holds the fp32 CUDA-core kernel to (log-mel within 1e-4 relative; bands more than 25 nats below
the loudest band of their frame skipped), plus what is specific to this kernel: the per-frame power-of-two scale (very
quiet and very loud input), tiles of four frames that straddle clips / channels, and agreement with the fp32 kernel."""
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


def tc(feat, x, **kw):
    return feat.mbe_device(x, kernel="tc", **kw)


@pytest.mark.parametrize("kind", ["mix", "noise", "chirp"])
@pytest.mark.parametrize("n", [1, 1000, 1023, 1024, 1025, 2048, 2049, 4096, 44100, 2 * 44100 + 1, 1024 * 37])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_tc_parity_mono(feat, kind, n, pad_mode):
    y = L.synth_clip(n % 97, n, 1, kind)[0]
    want = L.mbe(y, pad_mode=pad_mode)
    got = tc(feat, torch.from_numpy(y).cuda(), pad_mode=pad_mode).cpu().numpy()
    assert close(got, want) <= RTOL


@pytest.mark.parametrize("n_ch,n", [(2, 30001), (2, 30000), (3, 5000), (1, 3 * 1024)])
def test_tc_multichannel_batch_layout(feat, n_ch, n):
    """frames per (clip, channel) not a multiple of four: tiles straddle channels and clips; odd n: unaligned rows"""
    clips = np.stack([L.synth_clip(20 + i, n, n_ch, "mix") for i in range(3)])
    got = tc(feat, torch.from_numpy(clips).cuda()).cpu().numpy()
    assert got.shape == (3, 1 + n // 1024, n_ch * 40)
    for i in range(3):
        assert close(got[i], L.mbe_multichannel(clips[i])) <= RTOL


def test_tc_golden_fixtures(feat, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_oracle.npz"))
    for name in ("mix_1s", "noise_odd", "chirp_stereo", "short"):
        for pm in ("constant", "reflect"):
            got = tc(feat, torch.from_numpy(g[name + "_pcm"]).cuda(), pad_mode=pm).cpu().numpy()
            assert close(got, g[f"{name}_{pm}"]) <= RTOL, (name, pm)


@pytest.mark.parametrize("amp", [1e-6, 1e-3, 1.0, 3e4])
def test_tc_frame_scale_keeps_fp16_in_range(feat, amp):
    """fp16 operand planes: every frame is scaled by a power of two before the split and un-scaled after the log"""
    y = (L.synth_clip(5, 20000, 1, "mix")[0].astype(np.float64) * amp).astype(np.float32)
    got = tc(feat, torch.from_numpy(y).cuda()).cpu().numpy()
    assert close(got, L.mbe(y)) <= RTOL


def test_tc_agrees_with_fp32_kernel(feat):
    y = torch.from_numpy(L.synth_clip(9, 5 * 44100 + 17, 2, "mix")).cuda()
    a = feat.mbe_device(y, kernel="fp32").cpu().numpy().astype(np.float64)
    b = tc(feat, y).cpu().numpy().astype(np.float64)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1.0)) <= 2e-5


def test_tc_other_sample_rate_silence_determinism(feat):
    y = L.synth_clip(5, 22050, 1, "noise")[0]
    got = tc(feat, torch.from_numpy(y).cuda(), sr=22050).cpu().numpy()
    assert close(got, L.mbe(y, sr=22050)) <= RTOL
    z = tc(feat, torch.zeros(5000, device="cuda")).cpu().numpy()
    assert np.all(np.isneginf(z))
    d = torch.from_numpy(L.synth_clip(1, 400000, 2, "mix")).cuda()
    a = tc(feat, d)
    for _ in range(3):
        assert torch.equal(a, tc(feat, d))


@pytest.mark.parametrize("n_ch,n", [(1, 1), (1, 2047), (2, 30001), (2, 1024 * 9)])
def test_tc_int16_ingest_is_bit_identical_to_float(feat, n_ch, n):
    rng = np.random.default_rng(7 + n)
    pcm16 = rng.integers(-32768, 32768, size=(n_ch, n), dtype=np.int16)
    pcm16[pcm16 == 0] = 1
    as_float = pcm16.astype(np.float32) / np.float32(32768.0)
    a = tc(feat, torch.from_numpy(pcm16).cuda())
    b = tc(feat, torch.from_numpy(as_float).cuda())
    assert torch.equal(a, b)


def test_tc_full_size_properties(feat):
    """BASELINE size (3-min stereo clip): frame-shift invariance and the oracle on a strided sample of frames"""
    n = 180 * 44100
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal((2, n + 1024))).astype(np.float32)
    d = torch.from_numpy(y).cuda()
    a = tc(feat, d[:, 1024:].contiguous())
    b = tc(feat, d[:, :n].contiguous())
    assert a.shape == (7752, 80)
    assert torch.equal(a[1:-2], b[2:-1])
    idx = np.arange(0, 7752, 517)
    want = L.mbe(y[1, 1024:])[idx]
    assert close(a[:, 40:].cpu().numpy()[idx], want) <= RTOL
