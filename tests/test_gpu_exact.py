"""GPU parity tests of the EXACT engine (prib_acc_params.mode = 2, csrc/acc_exact.cu): the reference's own
log-domain arithmetic on the GPU.  The bar is bit-identity — every float of every vector equal, as uint32 —
with the unmodified reference compiled here (the committed fixtures of tests/golden) and with the pinned
oracle on fresh inputs, including the Z > 690 log-sum path, the Q1 float-overflow clamp regime, unknown
bases, lower case, the Q3/Q4 low-complexity construct and the shortest sequences."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

_ctx_cache = {}


def rac_exact(W, delta, **kw):
    from priblast_b200 import Raccess
    kw.setdefault("max_batch_bytes", 6 << 30)
    key = (W, delta, tuple(sorted(kw.items())))
    if key not in _ctx_cache:
        _ctx_cache[key] = Raccess(W, delta, mode=2, **kw)
    return _ctx_cache[key]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _assert_bits(got, want, what):
    g, w = _bits(got), _bits(want)
    assert g.shape == w.shape, what
    bad = np.nonzero(g != w)[0]
    assert bad.size == 0, (f"{what}: {bad.size}/{g.size} floats differ, first at {bad[0]}: "
                           f"got {np.asarray(got)[bad[0]]!r} want {np.asarray(want)[bad[0]]!r}")


def _groups():
    g = {}
    for c in GOLDEN:
        g.setdefault((c["W"], c["delta"]), []).append(c)
    return sorted(g.items())


@pytest.mark.parametrize("key,cases", _groups(), ids=lambda x: f"W{x[0]}_d{x[1]}" if isinstance(x, tuple) else "")
def test_golden_fixtures_bit_identical(key, cases):
    """All committed reference vectors, one batched C-ABI call per (W, delta): equal bit for bit."""
    W, delta = key
    res = rac_exact(W, delta).run_batch([c["seq"] for c in cases])
    for c, (acc, cond) in zip(cases, res):
        _assert_bits(acc, c["acc"], c["name"] + " acc")
        _assert_bits(cond, c["cond"], c["name"] + " cond")


def test_fresh_inputs_bit_identical_to_oracle(oracle_lib):
    rng = np.random.default_rng(20261018)
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in (0, 1, 3, 4, 5, 6, 7, 9, 12, 33, 64, 129, 400, 777)]
    seqs.append("".join("ACGUNacgut"[k] for k in rng.integers(0, 10, 300)))
    seqs.append("GGGGGCCCCC" * 20 + "AAAAAAA" * 10 + "GGGGGAAAACCCCC" * 15)  # low complexity (Q3)
    seqs.append("".join("ACGU"[k] for k in rng.integers(0, 4, 3300)))         # Z > 690: log-sum path (Q4 code)
    res = rac_exact(70, 5).run_batch(seqs)
    ora, _ = oracle_lib.run_batch(seqs, 70, 5)
    for s, (a, c), (oa, oc) in zip(seqs, res, ora):
        _assert_bits(a, oa, f"L={len(s)} acc")
        _assert_bits(c, oc, f"L={len(s)} cond")


@pytest.mark.parametrize("W,delta", [(20, 2), (40, 10), (150, 5)])
def test_span_and_window_sweep_bit_identical(oracle_lib, W, delta):
    rng = np.random.default_rng(W * 100 + delta)
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in (delta, delta + 1, 50, 222, 600)]
    res = rac_exact(W, delta).run_batch(seqs)
    ora, _ = oracle_lib.run_batch(seqs, W, delta)
    for s, (a, c), (oa, oc) in zip(seqs, res, ora):
        _assert_bits(a, oa, f"W={W} d={delta} L={len(s)} acc")
        _assert_bits(c, oc, f"W={W} d={delta} L={len(s)} cond")


def test_exact_is_independent_of_batching():
    from priblast_b200 import workloads
    seqs = workloads.cfg2(first=24)
    big = rac_exact(70, 5).run_batch(seqs)
    small = rac_exact(70, 5, max_batch_bytes=100 << 20)
    parts = small.run_batch(seqs)
    assert small.counters()["batches"] > 1
    for (a, c), (b, d) in zip(big, parts):
        assert np.array_equal(_bits(a), _bits(b)) and np.array_equal(_bits(c), _bits(d))


def test_exact_and_fast_engines_agree_within_reference_noise():
    """The two engines are independent formulations of the same recurrences: they must agree to the
    tolerance the fast engine is held to against the reference."""
    from conftest import ATOL_VS_REF, RTOL_VS_REF, assert_close_kcal
    from priblast_b200 import Raccess, workloads
    seqs = workloads.cfg2(first=16)
    ex = rac_exact(70, 5).run_batch(seqs)
    with Raccess(70, 5, max_batch_bytes=4 << 30) as fast:
        fa = fast.run_batch(seqs)
    for k, ((a, c), (b, d)) in enumerate(zip(ex, fa)):
        assert_close_kcal(b, a, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} acc")
        assert_close_kcal(d, c, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} cond")


def test_full_size_bench_batch_cross_checked_by_the_exact_engine():
    """BASELINE config 2 at the size bench.py times (1,536 transcripts, 2.8 M nt, one C-ABI call, one device
    batch: the tiling / halo / layout paths of a full batch): every output finite and zero-padded as the
    reference pads, and a strided sample of the batch recomputed by the exact engine — i.e. with the
    reference's own bits — within the fast engine's stated tolerance."""
    from conftest import ATOL_VS_REF, RTOL_VS_REF, assert_close_kcal
    from priblast_b200 import Raccess, workloads
    seqs = workloads.cfg2(first=1536)
    with Raccess(70, 5) as fast:
        res = fast.run_batch(seqs)
        assert fast.counters()["batches"] == 1
    for (a, c), s in zip(res, seqs):
        L = len(s)
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(c))
        assert np.all(a[L - 4:] == 0) and np.all(c[:5] == 0)
        assert a[:L - 4].min() > -7.0 and a[:L - 4].max() < 54.3  # Q4 floor / fmath::log(0f) ceiling
    idx = list(range(7, 1536, 96))
    ex = rac_exact(70, 5).run_batch([seqs[k] for k in idx])
    worst = 0.0
    for k, (ea, ec) in zip(idx, ex):
        worst = max(worst, assert_close_kcal(res[k][0], ea, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} acc"))
        worst = max(worst, assert_close_kcal(res[k][1], ec, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} cond"))
    print(f"full bench batch: {len(idx)} transcripts re-done with the reference's bits, max |d| = {worst:.2e} kcal/mol")


def test_long_lncrna_fast_engine_against_exact_engine():
    """BASELINE config 3 in miniature: a 25 kb lncRNA (Z ~ 6,500: far beyond the direct path), fast engine
    against the reference's bits from the exact engine."""
    from conftest import ATOL_VS_REF, RTOL_VS_REF, assert_close_kcal
    from priblast_b200 import Raccess, workloads
    seq = workloads.cfg3(first=2)[1][:25000]
    (ea, ec), = rac_exact(70, 5).run_batch([seq])
    with Raccess(70, 5, max_batch_bytes=4 << 30) as fast:
        a, c = fast.run(seq)
    assert_close_kcal(a, ea, ATOL_VS_REF, RTOL_VS_REF, "acc")
    assert_close_kcal(c, ec, ATOL_VS_REF, RTOL_VS_REF, "cond")
