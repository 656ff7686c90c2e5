"""The C-ABI library builds, loads and exports every symbol include/sedb200.h declares (no GPU needed:
nothing here launches a kernel)."""
import ctypes

import numpy as np

from oracle import logmel_ref as L
from sed_crnn_b200 import _lib


def test_exports_match_header(built_lib):
    declared = _lib.header_symbols()
    assert declared, "no symbols parsed from include/sedb200.h"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in sedb200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_frames(built_lib):
    assert built_lib.sedb200_version() == 100
    for n in (1, 1000, 1023, 1024, 2048, 44100, 7938000):
        assert built_lib.sedb200_logmel_frames(n) == 1 + n // 1024 == L.n_frames(n)
    assert built_lib.sedb200_logmel_frames(0) == 0


def test_host_side_mel_table_matches_oracle(built_lib):
    out = np.empty((40, 1025), np.float32)
    assert built_lib.sedb200_mel_filterbank(44100, out.ctypes.data) == 0
    ref = L.mel_filterbank()
    assert np.abs(out - ref).max() <= 1.2e-9 and int((out != 0).sum()) == 1945
    out2 = np.empty((40, 1025), np.float32)
    assert built_lib.sedb200_mel_filterbank(22050, out2.ctypes.data) == 0
    assert np.abs(out2 - L.mel_filterbank(sr=22050)).max() <= 2.5e-9


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    rc = built_lib.sedb200_logmel_f32(None, 1, 1, 0, 44100, 0, None, None)
    assert rc == _lib.EINVAL and b"empty" in built_lib.sedb200_last_error()
    rc = built_lib.sedb200_logmel_f32(None, 1, 1, 100, 44100, 7, None, None)
    assert rc == _lib.EINVAL
    assert built_lib.sedb200_logmel_host_scratch(2, 2, 4096) == 2 * 2 * 4096 * 4 + 2 * 5 * 80 * 4


def test_crnn_geometry_helpers_without_a_gpu(built_lib):
    """Host-only plan functions: every preset validates, the workspace grows with the batch, the parameter layout is
    256-byte aligned and counts the reference's tensors; the int16 log-mel scratch is sized for 2-byte samples."""
    from sed_crnn_b200 import config
    for name, cfg in config.PRESETS.items():
        d = cfg.desc()
        assert built_lib.sedb200_crnn_validate(ctypes.byref(d)) == 0, name
        assert built_lib.sedb200_crnn_seq_len(ctypes.byref(d)) == cfg.seq_len_out, name
        assert built_lib.sedb200_crnn_flat(ctypes.byref(d)) == cfg.flat, name
        n = built_lib.sedb200_crnn_n_tensors(ctypes.byref(d))
        assert n == 4 * len(cfg.pool) + 4 * len(cfg.gru_units) + 2 * (len(cfg.dense_units) + 1), name
        offs = (ctypes.c_long * n)()
        total = built_lib.sedb200_crnn_param_layout(ctypes.byref(d), offs)
        assert total == cfg.n_param_floats() and all(o % 64 == 0 for o in offs) and list(offs) == sorted(offs), name
        w8, w16 = (built_lib.sedb200_crnn_workspace_bytes(ctypes.byref(d), b) for b in (8, 16))
        assert 0 < w8 < w16, name
    assert built_lib.sedb200_logmel_host_scratch_i16(2, 2, 4096) == 2 * 2 * 4096 * 2 + 2 * 5 * 80 * 4
    rc = built_lib.sedb200_logmel_i16(None, 1, 1, 0, 44100, 0, None, None)
    assert rc == _lib.EINVAL
