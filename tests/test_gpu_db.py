"""End-to-end (BASELINE config 5 in miniature): database built by the GPU front-end, searched by the
UNMODIFIED reference `ris`; the hit list must be the one the all-CPU reference pipeline produces.
`.acc` differs from the reference only within the stated tolerance; all other files byte-for-byte."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ATOL_VS_REF, ROOT, RTOL_VS_REF, assert_close_kcal

pytestmark = pytest.mark.gpu

FRONT = os.path.join(ROOT, "priblast_b200", "pRIblast_b200")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "pRIblast_ref")


def _write(path, seqs, prefix):
    with open(path, "w") as f:
        for k, s in enumerate(seqs):
            f.write(f">{prefix}{k}\n")
            for p in range(0, len(s), 60):
                f.write(s[p:p + 60] + "\n")


def _read_acc(path, n):
    raw = open(path, "rb").read()
    out, p = [], 0
    for _ in range(n):
        n1 = int(np.frombuffer(raw[p:p + 4], np.int32)[0]); p += 4
        a = np.frombuffer(raw[p:p + 4 * n1], np.float32); p += 4 * n1
        L = int(np.frombuffer(raw[p:p + 4], np.int32)[0]); p += 4
        c = np.frombuffer(raw[p:p + 4 * L], np.float32); p += 4 * L
        out.append((a, c))
    assert p == len(raw)
    return out


def _hits(path):
    rows = []
    for ln in open(path):
        f = ln.rstrip("\n").split(",")
        # Id, Query name, Query Length, Target name, Target Length, Accessibility E, Hybridization E, Interaction E, BasePair
        if len(f) < 9:
            continue
        try:
            rows.append((f[1], f[3], f[8].strip(), float(f[5]), float(f[6]), float(f[7])))
        except ValueError:  # header lines of the ris output
            continue
    return sorted(rows)  # the reference merges per-thread files: line order varies from run to run


def test_gpu_db_then_reference_ris(tmp_path):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/pRIblast_ref not built")
    subprocess.run(["make", "-C", os.path.join(ROOT, "priblast_b200", "csrc", "host")], check=True,
                   stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(77)
    db_seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, int(L))) for L in rng.integers(150, 700, 36)]
    # plant reverse-complement targets so that ris finds strong interactions
    comp = {"A": "U", "C": "G", "G": "C", "U": "A"}
    queries = []
    for k in range(6):
        src = db_seqs[5 * k]
        st = int(rng.integers(20, len(src) - 60))
        site = "".join(comp[b] for b in reversed(src[st:st + 28]))
        pad = "".join("ACGU"[i] for i in rng.integers(0, 4, 160))
        queries.append(pad[:80] + site + pad[80:])
    fa, qa = str(tmp_path / "db.fa"), str(tmp_path / "q.fa")
    _write(fa, db_seqs, "t")
    _write(qa, queries, "q")
    env = dict(os.environ, OMP_NUM_THREADS="8")
    subprocess.run([REFBIN, "db", "-i", fa, "-o", str(tmp_path / "ref"), "-c", "16"], check=True, env=env, cwd=tmp_path)
    subprocess.run([FRONT, "db", "-i", fa, "-o", str(tmp_path / "gpu"), "-c", "16"], check=True, env=env)
    for ext in (".seq", ".ind", ".nam", ".bas"):
        assert open(str(tmp_path / "ref") + ext, "rb").read() == open(str(tmp_path / "gpu") + ext, "rb").read(), ext
    assert os.path.getsize(str(tmp_path / "ref.acc")) == os.path.getsize(str(tmp_path / "gpu.acc"))
    worst = 0.0
    for (ra, rc), (ga, gc) in zip(_read_acc(str(tmp_path / "ref.acc"), 36), _read_acc(str(tmp_path / "gpu.acc"), 36)):
        worst = max(worst, assert_close_kcal(ga, ra, ATOL_VS_REF, RTOL_VS_REF, "acc"))
        worst = max(worst, assert_close_kcal(gc, rc, ATOL_VS_REF, RTOL_VS_REF, "cond"))
    print(f".acc max |d| vs reference db: {worst:.2e} kcal/mol")
    # the unchanged reference `ris` on both databases
    for name in ("ref", "gpu"):
        subprocess.run([REFBIN, "ris", "-i", qa, "-o", str(tmp_path / f"hits_{name}.txt"), "-d", str(tmp_path / name)],
                       check=True, env=env, cwd=tmp_path)
    hr, hg = _hits(str(tmp_path / "hits_ref.txt")), _hits(str(tmp_path / "hits_gpu.txt"))
    assert len(hr) > 0, "test set produced no hits"
    assert [h[:3] for h in hr] == [h[:3] for h in hg], "structural hit list (query, target, interval) differs"
    for a, b in zip(hr, hg):
        assert abs(a[3] - b[3]) <= 2e-4 and abs(a[4] - b[4]) <= 1e-9 and abs(a[5] - b[5]) <= 2e-4
    print(f"{len(hr)} hits identical in structure; energies within 2e-4 kcal/mol")


def test_exact_mode_database_and_ris_output_byte_identical(tmp_path):
    """`db -m exact` (csrc/acc_exact.cu: the reference's own arithmetic on the GPU): ALL five database files
    byte-identical to the reference program's, and the unmodified reference `ris` prints the same lines."""
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/pRIblast_ref not built")
    subprocess.run(["make", "-C", os.path.join(ROOT, "priblast_b200", "csrc", "host")], check=True,
                   stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(78)
    lens = list(rng.integers(120, 900, 30)) + [5, 6, 3100]  # shortest legal records and one Z > 690 (log-sum path)
    db_seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, int(L))) for L in lens]
    db_seqs[3] = db_seqs[3][:100] + "NNNNnnacgu" + db_seqs[3][110:]
    comp = {"A": "U", "C": "G", "G": "C", "U": "A"}
    queries = []
    for k in range(5):
        src = db_seqs[4 * k + 1]
        st = int(rng.integers(10, len(src) - 50))
        site = "".join(comp.get(b, "A") for b in reversed(src[st:st + 26]))
        pad = "".join("ACGU"[i] for i in rng.integers(0, 4, 140))
        queries.append(pad[:70] + site + pad[70:])
    fa, qa = str(tmp_path / "db.fa"), str(tmp_path / "q.fa")
    _write(fa, db_seqs, "t")
    _write(qa, queries, "q")
    env = dict(os.environ, OMP_NUM_THREADS="8")
    subprocess.run([REFBIN, "db", "-i", fa, "-o", str(tmp_path / "ref"), "-c", "12"], check=True, env=env, cwd=tmp_path)
    subprocess.run([FRONT, "db", "-i", fa, "-o", str(tmp_path / "gpu"), "-c", "12", "-m", "exact"], check=True, env=env)
    for ext in (".acc", ".seq", ".ind", ".nam", ".bas"):
        assert open(str(tmp_path / "ref") + ext, "rb").read() == open(str(tmp_path / "gpu") + ext, "rb").read(), ext
    for name in ("ref", "gpu"):
        subprocess.run([REFBIN, "ris", "-i", qa, "-o", str(tmp_path / f"hits_{name}.txt"), "-d", str(tmp_path / name)],
                       check=True, env=env, cwd=tmp_path)
    def lines(fn):  # the reference numbers hits in the order its OpenMP threads finish: drop the Id column, sort
        return sorted(ln.split(",", 1)[1] if ln[:1].isdigit() else ln for ln in open(fn).read().splitlines()
                      if not ln.startswith("input:"))  # that header line echoes the database path

    lr, lg = lines(str(tmp_path / "hits_ref.txt")), lines(str(tmp_path / "hits_gpu.txt"))
    assert len(lr) > 2, "test set produced no hits"
    assert lr == lg, "ris output differs"  # every printed digit of every energy equal
