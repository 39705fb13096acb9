"""CPU tests of the checker itself: the plain-C restatement must be BIT-identical to the compiled
reference on the committed fixtures (and on fresh inputs when oracle/_ref is present)."""
import os
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("case", [c for c in GOLDEN if len(c["seq"]) <= 600], ids=lambda c: c["name"])
def test_restatement_matches_golden_bits(oracle_lib, case):
    acc, cond = oracle_lib.run(case["seq"], case["W"], case["delta"])
    assert np.array_equal(_bits(acc), _bits(case["acc"]))
    assert np.array_equal(_bits(cond), _bits(case["cond"]))


@pytest.mark.parametrize("name", ["rand_L2900_W70_d5", "gcstem_polyA_L3360"])
def test_restatement_matches_golden_bits_logpath(oracle_lib, name):
    case = next(c for c in GOLDEN if c["name"] == name)  # Z > 690: log-sum biloop path, Q4
    acc, cond = oracle_lib.run(case["seq"], case["W"], case["delta"])
    assert np.array_equal(_bits(acc), _bits(case["acc"]))
    assert np.array_equal(_bits(cond), _bits(case["cond"]))


def test_restatement_matches_reference_on_fresh_inputs(oracle_lib, ref_lib):
    rng = np.random.default_rng(20261018)
    for L, W, delta in [(37, 70, 5), (211, 20, 3), (333, 70, 7), (150, 150, 2), (420, 70, 5)]:
        seq = "".join("ACGUNacgt"[k] for k in rng.integers(0, 9, L))
        a, c = oracle_lib.run(seq, W, delta)
        ra, rc = ref_lib.run(seq, W, delta)
        assert np.array_equal(_bits(a), _bits(ra)) and np.array_equal(_bits(c), _bits(rc)), (L, W, delta)


def test_reference_file_record_equals_vector_form(ref_lib):
    """Run(seq, idx) (raccess.cpp:447-481) writes: n1, acc[0:n1], L, cond[0:L]."""
    case = next(c for c in GOLDEN if c["name"] == "rand_L100_W70_d5")
    with tempfile.TemporaryDirectory() as d:
        data = ref_lib.run_file_bytes(case["seq"], case["W"], case["delta"], d, 3)
    L, delta = 100, 5
    n1 = np.frombuffer(data[:4], dtype=np.int32)[0]
    assert n1 == L - delta + 1 and len(data) == 8 + 4 * (2 * L - delta + 1)
    acc = np.frombuffer(data[4:4 + 4 * n1], dtype=np.float32)
    n2 = np.frombuffer(data[4 + 4 * n1:8 + 4 * n1], dtype=np.int32)[0]
    cond = np.frombuffer(data[8 + 4 * n1:], dtype=np.float32)
    assert n2 == L
    assert np.array_equal(_bits(acc), _bits(case["acc"][:n1]))
    assert np.array_equal(_bits(cond), _bits(case["cond"]))


def test_reference_noise_vs_exact(oracle_lib):
    """Documents the reference's own approximation noise (float table log in every log-sum)."""
    case = next(c for c in GOLDEN if c["name"] == "rand_L360_W70_d5")
    ea, ec = oracle_lib.run_exact(case["seq"], case["W"], case["delta"])
    noise = max(np.abs(ea - case["acc"]).max(), np.abs(ec - case["cond"]).max())
    assert 1e-7 < noise < 1e-4


def test_term_counter(oracle_lib):
    case = next(c for c in GOLDEN if c["name"] == "rand_L500_W70_d5")
    cnt = oracle_lib.count_terms(case["seq"], 70, 5)
    per_nt = (cnt["lse_inside"] + cnt["lse_outside"] + cnt["lse_access"] + cnt["expd_access"]) / 500
    assert 8_000 < per_nt < 25_000  # ~16 k terms per nt at W=70


def test_fmath_edge_values(oracle_lib):
    import ctypes
    lib = oracle_lib.lib
    lib.fmr_init()
    lib.fmr_logf.restype = ctypes.c_float
    lib.fmr_logf.argtypes = [ctypes.c_float]
    lib.fmr_expd.restype = ctypes.c_double
    lib.fmr_expd.argtypes = [ctypes.c_double]
    assert lib.fmr_logf(1.0) == 0.0
    assert abs(lib.fmr_logf(0.0) + 88.0297) < 1e-3
    assert abs(lib.fmr_logf(float("inf")) - 88.7228) < 1e-3
    assert lib.fmr_expd(-708.4) == 0.0
    for x in (-0.5, -3.25, -20.0, 1.5):
        assert abs(lib.fmr_expd(x) / np.exp(x) - 1) < 1e-12
