"""GPU parity tests: the CUDA path, called through the C ABI (priblast_b200.Raccess -> libpriblast_acc.so),
against the committed reference fixtures, the oracle on fresh inputs, and size-independent properties."""
import numpy as np
import pytest

from conftest import ATOL_VS_EXACT, ATOL_VS_REF, GOLDEN, RTOL_VS_EXACT, RTOL_VS_REF, assert_close_kcal

pytestmark = pytest.mark.gpu

_ctx_cache = {}


def rac(W, delta, **kw):
    """One context per (W, delta); each gets a bounded DP budget so a dozen of them fit one GPU."""
    from priblast_b200 import Raccess
    kw.setdefault("max_batch_bytes", 8 << 30)
    key = (W, delta, tuple(sorted(kw.items())))
    if key not in _ctx_cache:
        _ctx_cache[key] = Raccess(W, delta, **kw)
    return _ctx_cache[key]


def _groups():
    g = {}
    for c in GOLDEN:
        g.setdefault((c["W"], c["delta"]), []).append(c)
    return sorted(g.items())


@pytest.mark.parametrize("key,cases", _groups(), ids=lambda x: f"W{x[0]}_d{x[1]}" if isinstance(x, tuple) else "")
def test_golden_fixtures(key, cases):
    """Every committed reference vector (all of tests/golden), one batched C-ABI call per (W, delta)."""
    W, delta = key
    res = rac(W, delta).run_batch([c["seq"] for c in cases])
    worst = 0.0
    for c, (acc, cond) in zip(cases, res):
        worst = max(worst, assert_close_kcal(acc, c["acc"], ATOL_VS_REF, RTOL_VS_REF, c["name"] + " acc"))
        worst = max(worst, assert_close_kcal(cond, c["cond"], ATOL_VS_REF, RTOL_VS_REF, c["name"] + " cond"))
        d = c["delta"]
        assert np.all(cond[:d] == 0) and np.all(acc[len(acc) - d + 1:] == 0), c["name"]
    print(f"W={W} delta={delta}: {len(cases)} cases, max |d| vs reference = {worst:.3e} kcal/mol")


def test_fresh_random_vs_oracle(oracle_lib):
    rng = np.random.default_rng(4242)
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in (17, 64, 129, 257, 400, 511, 777)]
    seqs.append("".join("ACGUNacgut"[k] for k in rng.integers(0, 10, 300)))
    res = rac(70, 5).run_batch(seqs)
    ora, _ = oracle_lib.run_batch(seqs, 70, 5)
    for s, (a, c), (oa, oc) in zip(seqs, res, ora):
        assert_close_kcal(a, oa, ATOL_VS_REF, RTOL_VS_REF, f"L={len(s)} acc")
        assert_close_kcal(c, oc, ATOL_VS_REF, RTOL_VS_REF, f"L={len(s)} cond")


def test_vs_exact_math_twin(oracle_lib):
    """Against exact libm the CUDA path is at float-rounding level (the reference is not)."""
    rng = np.random.default_rng(99)
    seq = "".join("ACGU"[k] for k in rng.integers(0, 4, 900))
    a, c = rac(70, 5).run(seq)
    ea, ec = oracle_lib.run_exact(seq, 70, 5)
    assert_close_kcal(a, ea, ATOL_VS_EXACT, RTOL_VS_EXACT, "acc")
    assert_close_kcal(c, ec, ATOL_VS_EXACT, RTOL_VS_EXACT, "cond")


def _cfg2_sample(n):
    from priblast_b200 import workloads
    return workloads.cfg2(first=n)


def test_batch_split_and_order_invariance():
    """Results must not depend on batching or input order (bit-identical)."""
    seqs = _cfg2_sample(48)
    big = rac(70, 5).run_batch(seqs)
    small = rac(70, 5, max_batch_bytes=400 << 20).run_batch(seqs)  # forces several device batches
    assert rac(70, 5, max_batch_bytes=400 << 20).counters()["batches"] > 1
    perm = np.random.default_rng(5).permutation(len(seqs))
    shuf = rac(70, 5).run_batch([seqs[k] for k in perm])
    for k in range(len(seqs)):
        assert np.array_equal(big[k][0].view(np.uint32), small[k][0].view(np.uint32))
        assert np.array_equal(big[k][1].view(np.uint32), small[k][1].view(np.uint32))
    for j, k in enumerate(perm):
        assert np.array_equal(big[k][0].view(np.uint32), shuf[j][0].view(np.uint32))
        assert np.array_equal(big[k][1].view(np.uint32), shuf[j][1].view(np.uint32))


def test_alphabet_equivalences():
    """raccess.cpp:55-68: case-insensitive, T == U, anything else is 'unknown'."""
    rng = np.random.default_rng(8)
    s = "".join("ACGU"[k] for k in rng.integers(0, 4, 350))
    r = rac(70, 5)
    base = r.run(s)
    for variant in (s.lower(), s.replace("U", "T"), s.replace("U", "t")):
        got = r.run(variant)
        assert np.array_equal(base[0].view(np.uint32), got[0].view(np.uint32))
        assert np.array_equal(base[1].view(np.uint32), got[1].view(np.uint32))
    a_n, _ = r.run(s[:100] + "N" + s[101:])
    a_x, _ = r.run(s[:100] + "X" + s[101:])
    assert np.array_equal(a_n.view(np.uint32), a_x.view(np.uint32))


def test_full_size_cfg1_sample_vs_reference(oracle_lib):
    """cfg1 (1,000 x 500 nt, W=70, delta=5: the Q1 clamp regime) — all 1,000 on the GPU, a slice checked
    against the compiled reference (or the restatement), the rest through properties."""
    from priblast_b200 import workloads
    from oracle_py import RefLib
    seqs = workloads.cfg1()
    res = rac(70, 5).run_batch(seqs)
    assert len(res) == 1000
    idx = list(range(0, 1000, 40))
    checker = RefLib() if RefLib.available() else oracle_lib
    ora, _ = checker.run_batch([seqs[k] for k in idx], 70, 5)
    for k, (oa, oc) in zip(idx, ora):
        assert_close_kcal(res[k][0], oa, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} acc")
        assert_close_kcal(res[k][1], oc, ATOL_VS_REF, RTOL_VS_REF, f"seq {k} cond")
    allv = np.concatenate([np.concatenate(r) for r in res])
    assert np.all(np.isfinite(allv))
    for a, c in res:
        assert np.all(a[:496] > -1.0) and np.all(a[:496] < 60.0)
        assert np.all(c[:5] == 0) and np.all(a[496:] == 0)


def test_long_sequence_log_path(oracle_lib):
    """Z > 690 (log-sum biloop path of raccess.cpp:683-771) on an 6 kb sequence."""
    from priblast_b200 import workloads
    seq = workloads.cfg3(first=1)[0][:6000]
    a, c = rac(70, 5).run(seq)
    ea, ec = oracle_lib.run_exact(seq, 70, 5)
    assert_close_kcal(a, ea, ATOL_VS_EXACT, RTOL_VS_EXACT, "acc")
    assert_close_kcal(c, ec, ATOL_VS_EXACT, RTOL_VS_EXACT, "cond")


def test_fp32_engine_flags_and_fp64_rerun():
    """The float engine must flag sequences whose band values leave its safe range (perfect GC hairpins:
    Boltzmann weights ~e^170) and the double engine must recompute exactly those, on the GPU."""
    names = ["perfect_hairpin_L70", "rand_L500_W70_d5", "gcstem_polyA_L576", "rand_L300_W70_d5"]
    cases = [next(c for c in GOLDEN if c["name"] == n) for n in names]
    r = rac(70, 5)
    before = r.counters()["fp64_rerun_sequences"]
    res = r.run_batch([c["seq"] for c in cases])
    rerun = r.counters()["fp64_rerun_sequences"] - before
    assert 1 <= rerun <= 2, rerun  # the hairpin (and possibly the GC-stem repeat), never the random ones
    for c, (a, cc) in zip(cases, res):
        assert_close_kcal(a, c["acc"], ATOL_VS_REF, RTOL_VS_REF, c["name"])
        assert_close_kcal(cc, c["cond"], ATOL_VS_REF, RTOL_VS_REF, c["name"])


def test_fp64_only_mode_matches_auto_mode():
    seqs = _cfg2_sample(12)
    auto = rac(70, 5).run_batch(seqs)
    f64 = rac(70, 5, mode=1).run_batch(seqs)
    assert rac(70, 5, mode=1).counters()["fp64_rerun_sequences"] == 0
    for (a, c), (b, d) in zip(auto, f64):
        assert_close_kcal(a, b, 5e-6, 5e-7, "acc fp32 vs fp64 engine")
        assert_close_kcal(c, d, 5e-6, 5e-7, "cond fp32 vs fp64 engine")


def test_edge_lengths():
    r = rac(70, 5)
    res = r.run_batch(["", "A", "ACG", "ACGU", "ACGUA", "GGGAAACCC"])
    assert [len(a) for a, _ in res] == [0, 1, 3, 4, 5, 9]
    for a, c in res[:4]:  # L < delta: the vector form returns zeros (raccess.cpp:510-527 never iterate)
        assert np.all(a == 0) and np.all(c == 0)
    assert np.all(res[4][0][1:] == 0)  # L == delta: one window, P = 1 -> -0.0 (raccess.cpp:515)


def test_staged_api_and_counters():
    seqs = _cfg2_sample(16)
    r = rac(70, 5)
    want = r.run_batch(seqs)
    nt = r.stage(seqs)
    r.compute()
    r.sync()
    got = r.fetch()
    for (a, c), (b, d) in zip(want, got):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(c.view(np.uint32), d.view(np.uint32))
    cnt = r.counters()
    assert cnt["nucleotides"] >= nt and cnt["kernel_ms"] > 0 and cnt["kernel_launches"] > 0


def test_pipelined_run_equals_the_split_calls_over_several_batches():
    """prib_acc_run fills its arena and launches batch by batch; prib_acc_stage + _compute + _fetch stage everything
    first.  Same bits either way, over several device batches and with a sequence the FP64 engine has to redo; a
    second prib_acc_compute on what the pipelined run staged must work too."""
    hairpin = next(c["seq"] for c in GOLDEN if c["name"] == "perfect_hairpin_L70")
    seqs = _cfg2_sample(40) + [hairpin.encode() if isinstance(hairpin, str) else hairpin] + _cfg2_sample(48)[40:]
    r = rac(70, 5, max_batch_bytes=400 << 20)
    c0 = r.counters()
    want = r.run_batch(seqs)
    c1 = r.counters()
    assert c1["batches"] - c0["batches"] > 2 and c1["fp64_rerun_sequences"] - c0["fp64_rerun_sequences"] >= 1
    r.compute()  # again, on the batches the run left staged
    again = r_fetch_after_run(r, seqs)
    r.stage(seqs)
    r.compute()
    got = r.fetch()
    for (a, c), (b, d), (e, f) in zip(want, got, again):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(c.view(np.uint32), d.view(np.uint32))
        assert np.array_equal(a.view(np.uint32), e.view(np.uint32)) and np.array_equal(c.view(np.uint32), f.view(np.uint32))


def r_fetch_after_run(r, seqs):
    """fetch() of the Python mirror needs the lengths of the staged set (run_batch does not record them)."""
    r._staged_lens = np.array([len(s) for s in seqs], np.int32)
    return r.fetch()


def test_record_bytes_match_reference_file():
    """`.acc` record of one sequence equals the fixture's record layout with GPU numbers inside."""
    case = next(c for c in GOLDEN if c["name"] == "rand_L100_W70_d5")
    r = rac(70, 5)
    acc, cond = r.run(case["seq"])
    rec = r.record_bytes(acc, cond)
    n1 = np.frombuffer(rec[:4], np.int32)[0]
    assert n1 == 96 and len(rec) == 8 + 4 * (200 - 5 + 1)
    assert_close_kcal(np.frombuffer(rec[4:4 + 4 * n1], np.float32), case["acc"][:n1], ATOL_VS_REF, RTOL_VS_REF)


@pytest.mark.parametrize("W,delta", [(1, 2), (3, 2), (4, 3), (5, 5), (6, 2), (7, 5), (10, 5), (12, 4), (33, 3), (34, 30),
                                      (99, 7), (200, 5)])
def test_span_and_window_sweep_vs_oracle(oracle_lib, W, delta):
    """Spans around every structural threshold of the kernels (no band cells, first stem at 5, time-tile group
    boundaries, MAXLOOP = 30 +- a few, the widest span) and window lengths up to 30, against the oracle."""
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in (40, 200, 333, 75)]
    with __import__("priblast_b200").Raccess(W, delta, max_batch_bytes=4 << 30) as r:
        got = r.run_batch(seqs)
    want, _ = oracle_lib.run_batch(seqs, W, delta)
    for s, (a, c), (oa, oc) in zip(seqs, got, want):
        assert_close_kcal(a, oa, ATOL_VS_REF, RTOL_VS_REF, f"W={W} d={delta} L={len(s)} acc")
        assert_close_kcal(c, oc, ATOL_VS_REF, RTOL_VS_REF, f"W={W} d={delta} L={len(s)} cond")


def test_repeated_runs_are_bit_identical():
    """The tile kernels synchronise warp to warp through progress counters (acc_tile.h, "Synchronisation"): a missed
    dependency would show up as run-to-run differences.  Twelve passes over the same 768 transcripts, two tile-kernel
    engines (FP32 and FP64), every output bit compared."""
    from priblast_b200 import Raccess
    seqs = _cfg2_sample(768)
    for mode in (0, 1):
        with Raccess(70, 5, mode=mode, max_batch_bytes=(24 if mode == 0 else 40) << 30) as r:
            r.stage(seqs if mode == 0 else seqs[:256])
            ref = None
            for it in range(12 if mode == 0 else 4):
                r.compute()
                out = r.fetch()
                flat = np.concatenate([np.concatenate([a, c]) for a, c in out]).view(np.uint32).copy()
                if ref is None:
                    ref = flat
                else:
                    assert np.array_equal(ref, flat), f"mode {mode}: pass {it} differs from pass 0 in {(ref != flat).sum()} values"
