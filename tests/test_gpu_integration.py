"""INTEGRATION.md's patch, compiled and run (VERDICT r1 "next" #8): the otherwise UNMODIFIED reference program with
only DbConstruction::CalculateAccessibility and ::ConstructSuffixArray swapped for oracle/integration/*.inc
(oracle/Makefile target `refpatched` -> oracle/_ref/pRIblast_patched, linked against libpriblast_acc.so) must
build the database the reference builds: byte for byte in exact mode, `.acc` within tolerance in fast mode."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ATOL_VS_REF, ROOT, RTOL_VS_REF, assert_close_kcal
from test_gpu_db import _read_acc, _write

pytestmark = pytest.mark.gpu

REFBIN = os.path.join(ROOT, "oracle", "_ref", "pRIblast_ref")
PATCHED = os.path.join(ROOT, "oracle", "_ref", "pRIblast_patched")
EXTS = ("bas", "nam", "acc", "seq", "ind")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    if not (os.path.exists(REFBIN) and os.path.exists(PATCHED)):
        pytest.skip("oracle/_ref/pRIblast_ref / pRIblast_patched not built (make -C oracle refbin refpatched)")
    d = tmp_path_factory.mktemp("integ")
    rng = np.random.default_rng(2024)
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, int(L))) for L in rng.integers(60, 900, 40)]
    seqs[3] = seqs[3].lower()
    seqs[7] = seqs[7][:100] + "NNNN" + seqs[7][104:]
    fa = str(d / "db.fa")
    _write(fa, seqs, "t")
    env = dict(os.environ, OMP_NUM_THREADS="4")

    def run(binary, name, mode=None, extra=()):
        e = dict(env)
        if mode:
            e["PRIB_ACC_MODE"] = mode
        subprocess.run([binary, "db", "-i", fa, "-o", str(d / name), "-p", str(d), *extra], check=True, env=e, cwd=str(d),
                       stdout=subprocess.DEVNULL)
        return {x: open(str(d / f"{name}.{x}"), "rb").read() for x in EXTS}

    return seqs, run


def test_patched_reference_exact_mode_is_byte_identical(built):
    seqs, run = built
    ref = run(REFBIN, "ref")
    got = run(PATCHED, "pat_exact", mode="exact")
    for x in EXTS:
        assert got[x] == ref[x], f"<db>.{x} differs from the unmodified reference's"


def test_patched_reference_fast_mode_and_paging(built, tmp_path):
    seqs, run = built
    ref = run(REFBIN, "ref_c", extra=("-c", "16", "-w", "40", "-d", "6"))
    got = run(PATCHED, "pat_fast", extra=("-c", "16", "-w", "40", "-d", "6"))
    for x in ("bas", "nam", "seq", "ind"):  # the suffix arrays come from prib_suffix_array here
        assert got[x] == ref[x], f"<db>.{x} differs from the unmodified reference's"
    pr, pg = str(tmp_path / "r.acc"), str(tmp_path / "g.acc")
    open(pr, "wb").write(ref["acc"])
    open(pg, "wb").write(got["acc"])
    for (ra, rc), (ga, gc) in zip(_read_acc(pr, len(seqs)), _read_acc(pg, len(seqs))):
        assert_close_kcal(ga, ra, ATOL_VS_REF, RTOL_VS_REF, "acc")
        assert_close_kcal(gc, rc, ATOL_VS_REF, RTOL_VS_REF, "cond")
