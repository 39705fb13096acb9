import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "raccess_golden.npz"))
    cases, off = [], 0
    for name, seq, W, delta, L in zip(g["names"], g["seqs"], g["W"], g["delta"], g["lens"]):
        L = int(L)
        cases.append(dict(name=str(name), seq=str(seq), W=int(W), delta=int(delta),
                          acc=g["acc"][off:off + L], cond=g["cond"][off:off + L]))
        off += L
    return cases


GOLDEN = load_golden()

# Tolerance of the fast (FP64 linear-domain) path against the REFERENCE, in kcal/mol.
# The reference evaluates every log-sum with a float-precision table log (raccess.cpp:414-419); against an
# exact-libm twin of itself it deviates by up to 1.7e-5 (L<=500) and 4.3e-5 (L~3000) on these fixtures
# (tests/test_oracle.py::test_reference_noise_vs_exact), while the fast path stays within 5e-6 of exact
# math.  |fast - reference| is therefore bounded by the reference's own noise; 1e-4 + 1e-6*|x| covers it.
ATOL_VS_REF = 1e-4
RTOL_VS_REF = 1e-6
# The mean is gated as well (SURVEY §7 asked for mean <= 2e-6 against exact math; against the REFERENCE the mean is
# bounded by the reference's own float-log noise, ~2e-6): a systematic offset would pass a max-only gate.
MEAN_VS_REF = 5e-6
ATOL_VS_EXACT = 6e-6
RTOL_VS_EXACT = 3e-7


def assert_close_kcal(got, want, atol, rtol, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, what
    if got.size == 0:
        return 0.0
    assert np.all(np.isfinite(got)), f"{what}: non-finite output"
    err = np.abs(got - want)
    lim = atol + rtol * np.abs(want)
    k = int(np.argmax(err - lim))
    assert err[k] <= lim[k], f"{what}: |d|={err[k]:.3e} kcal/mol at {k} (got {got[k]!r}, want {want[k]!r})"
    return float(err.max())


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle_py import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    from oracle_py import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return RefLib()


def parity_stats(got, want):
    """max/mean absolute and max relative deviation (kcal/mol) of two lists of (acc, cond) results."""
    g = np.concatenate([np.concatenate([np.asarray(a, np.float64), np.asarray(c, np.float64)]) for a, c in got])
    w = np.concatenate([np.concatenate([np.asarray(a, np.float64), np.asarray(c, np.float64)]) for a, c in want])
    assert g.shape == w.shape
    assert np.all(np.isfinite(g)), "non-finite output"
    err = np.abs(g - w)
    nz = np.abs(w) > 1e-3
    return {"n": int(g.size), "max_abs": float(err.max()), "mean_abs": float(err.mean()),
            "max_rel": float((err[nz] / np.abs(w[nz])).max()) if nz.any() else 0.0}
