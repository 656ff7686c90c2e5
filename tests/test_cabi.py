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
