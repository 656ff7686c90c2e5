"""Pin oracle/metrics_ref.py bit-exactly to the reference's metrics.py (golden KATs were produced by
executing /root/reference/metrics.py -- oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import metrics_ref as M
from oracle import ref_import


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_kat.npz"))
    for n in g["names"]:
        yield str(n), g[f"{n}_O"], g[f"{n}_T"], int(g[f"{n}_block"]), g[f"{n}_out"]


def test_metrics_restatement_is_bit_exact(golden_dir):
    for name, O, T, block, want in _cases(golden_dir):
        got = np.array([M.f1_overall_framewise(O, T), M.er_overall_framewise(O, T),
                        M.f1_overall_1sec(O, T, block), M.er_overall_1sec(O, T, block)])
        assert np.array_equal(got, want, equal_nan=True), (name, got, want)


def test_survey_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_kat.npz"))
    assert g["m1_out"].tolist() == [0.33333333333333315, 1.7142857142857142, 1.0, 0.0]
    assert g["m2_out"].tolist() == [0.22616195495927158, 0.9619140625, 0.7999999999999999, 0.5]
    assert g["no_ref_out"][0] == 0.0 and np.isinf(g["no_ref_out"][1])


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")
def test_against_live_reference_random():
    ref = ref_import.load("metrics")
    rng = np.random.default_rng(0)
    for _ in range(20):
        n, t, c = rng.integers(1, 9), rng.integers(1, 40), rng.integers(1, 7)
        O = (rng.random((n, t, c)) > rng.random()).astype(np.uint8)
        T = (rng.random((n, t, c)) > rng.random()).astype(np.float32)
        blk = int(rng.integers(1, 12))
        with np.errstate(all="ignore"):
            want = [ref.f1_overall_framewise(O, T), ref.er_overall_framewise(O, T),
                    ref.f1_overall_1sec(O, T, blk), ref.er_overall_1sec(O, T, blk)]
        got = [M.f1_overall_framewise(O, T), M.er_overall_framewise(O, T),
               M.f1_overall_1sec(O, T, blk), M.er_overall_1sec(O, T, blk)]
        assert np.array_equal(np.array(got), np.array(want), equal_nan=True)
