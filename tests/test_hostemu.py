"""CPU check of the restructured mathematics: the per-thread bodies of the CUDA kernels
(priblast_b200/csrc/acc_core.h), run in plain loops by tests/hostemu, against the golden fixtures.
This validates the formulation where no GPU exists; the `-m gpu` tests validate the real kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ATOL_VS_EXACT, ATOL_VS_REF, GOLDEN, ROOT, RTOL_VS_EXACT, RTOL_VS_REF, assert_close_kcal

_f32p = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="session")
def emu():
    d = os.path.join(ROOT, "tests", "hostemu")
    subprocess.run(["make", "-C", d], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(d, "libhostemu.so"))
    lib.hostemu_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p]

    def run(seq, W, delta):
        L = len(seq)
        a = np.zeros(max(L, 1), np.float32)
        c = np.zeros(max(L, 1), np.float32)
        assert lib.hostemu_run(seq.encode(), L, W, delta, a.ctypes.data_as(_f32p), c.ctypes.data_as(_f32p)) == 0
        return a[:L], c[:L]

    return run


@pytest.mark.parametrize("case", [c for c in GOLDEN if len(c["seq"]) <= 1600], ids=lambda c: c["name"])
def test_formulation_vs_reference(emu, case):
    acc, cond = emu(case["seq"], case["W"], case["delta"])
    assert_close_kcal(acc, case["acc"], ATOL_VS_REF, RTOL_VS_REF, "acc")
    assert_close_kcal(cond, case["cond"], ATOL_VS_REF, RTOL_VS_REF, "cond")
    d = case["delta"]
    assert np.all(cond[:d] == 0) and np.all(acc[len(acc) - d + 1:] == 0)


@pytest.mark.parametrize("name", ["rand_L500_W70_d5", "rand_L300_W150_d2", "gcstem_polyA_L576",
                                  "perfect_hairpin_L150_W150", "mixed_case_N_L300"])
def test_formulation_vs_exact_math(emu, oracle_lib, name):
    """Against the exact-libm twin the fast formulation is at float-rounding level."""
    case = next(c for c in GOLDEN if c["name"] == name)
    ea, ec = oracle_lib.run_exact(case["seq"], case["W"], case["delta"])
    acc, cond = emu(case["seq"], case["W"], case["delta"])
    assert_close_kcal(acc, ea, ATOL_VS_EXACT, RTOL_VS_EXACT, "acc")
    assert_close_kcal(cond, ec, ATOL_VS_EXACT, RTOL_VS_EXACT, "cond")
