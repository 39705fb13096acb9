"""BASELINE config 5 at a meaningful size (VERDICT r1 "next" #9): a database of N cfg2 transcripts built three ways —
the unmodified reference `db` on the host CPUs, `pRIblast_b200 db` with the fast engine, and with the exact engine — then
the UNCHANGED reference `ris` with Q cfg5 queries on all three; the hit lists are diffed structurally (query, target,
base-pair intervals) and textually (every printed digit).

  python profiles/cfg5_run.py [n_transcripts=2000] [n_queries=50] [out.json]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from priblast_b200 import workloads

REFBIN = os.path.join(ROOT, "oracle", "_ref", "pRIblast_ref")
FRONT = os.path.join(ROOT, "priblast_b200", "pRIblast_b200")


def cfg5_queries(n, db_seqs, seed=7):
    """cfg5: lengths clip(LogNormal(ln 1000, 0.6), 200, 5000), seed 7.  Every query carries the reverse
    complement of a 30-nt site of one database transcript, so that each has at least one strong interaction."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(rng.lognormal(np.log(1000.0), 0.6, n)), 200, 5000).astype(int)
    comp = {65: 85, 67: 71, 71: 67, 85: 65}
    out = []
    for k, L in enumerate(lens):
        q = bytearray(np.frombuffer(b"ACGU", np.uint8)[rng.integers(0, 4, L)].tobytes())
        src = db_seqs[int(rng.integers(0, len(db_seqs)))]
        st = int(rng.integers(0, len(src) - 30))
        site = bytes(comp[b] for b in reversed(src[st:st + 30]))
        p = int(rng.integers(0, L - 30))
        q[p:p + 30] = site
        out.append(bytes(q))
    return out


def hits(path):
    rows, text = [], []
    for ln in open(path):
        if ln.startswith("input:"):
            continue
        f = ln.rstrip("\n").split(",")
        body = ln.split(",", 1)[1] if ln[:1].isdigit() else ln  # drop the Id column (thread-order dependent)
        text.append(body)
        if len(f) >= 9:
            try:
                rows.append((f[1], f[3], f[8].strip(), float(f[5]), float(f[6]), float(f[7])))
            except ValueError:
                pass
    return sorted(rows), sorted(text)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    db_seqs = workloads.cfg2(first=n)
    queries = cfg5_queries(nq, db_seqs)
    res = {"transcripts": n, "nt": sum(map(len, db_seqs)), "queries": nq, "query_nt": sum(map(len, queries)),
           "host_cores": len(os.sched_getaffinity(0))}
    with tempfile.TemporaryDirectory() as d:
        fa, qa = os.path.join(d, "db.fa"), os.path.join(d, "q.fa")
        workloads.write_fasta(fa, db_seqs, prefix="t")
        workloads.write_fasta(qa, queries, prefix="q")
        env = dict(os.environ)
        t0 = time.perf_counter()
        subprocess.run([REFBIN, "db", "-i", fa, "-o", os.path.join(d, "ref"), "-p", d], check=True, env=env, cwd=d)
        res["reference_db_seconds"] = round(time.perf_counter() - t0, 2)
        for name, mode in (("fast", "auto"), ("exact", "exact")):
            t0 = time.perf_counter()
            subprocess.run([FRONT, "db", "-i", fa, "-o", os.path.join(d, name), "-m", mode], check=True, env=env)
            res[f"gpu_db_{name}_seconds"] = round(time.perf_counter() - t0, 2)
        for ext in ("seq", "ind", "nam", "bas", "acc"):
            ref = open(os.path.join(d, "ref." + ext), "rb").read()
            res[f"{ext}_identical_fast"] = open(os.path.join(d, "fast." + ext), "rb").read() == ref
            res[f"{ext}_identical_exact"] = open(os.path.join(d, "exact." + ext), "rb").read() == ref
        out = {}
        for name in ("ref", "fast", "exact"):
            t0 = time.perf_counter()
            subprocess.run([REFBIN, "ris", "-i", qa, "-o", os.path.join(d, f"hits_{name}.txt"), "-d", os.path.join(d, name)],
                           check=True, env=env, cwd=d)
            res[f"ris_on_{name}_db_seconds"] = round(time.perf_counter() - t0, 2)
            out[name] = hits(os.path.join(d, f"hits_{name}.txt"))
        rr, rt = out["ref"]
        res["hits_reference"] = len(rr)
        for name in ("fast", "exact"):
            gr, gt = out[name]
            same_struct = [a[:3] for a in rr] == [b[:3] for b in gr]
            res[f"hits_{name}"] = len(gr)
            res[f"structural_hit_list_identical_{name}"] = same_struct
            res[f"text_identical_{name}"] = rt == gt
            if same_struct and rr:
                res[f"max_energy_difference_{name}"] = max(max(abs(a[3] - b[3]), abs(a[4] - b[4]), abs(a[5] - b[5]))
                                                           for a, b in zip(rr, gr))
            res[f"printed_lines_differing_{name}"] = sum(1 for a, b in zip(rt, gt) if a != b) + abs(len(rt) - len(gt))
    s = json.dumps(res, indent=1)
    print(s)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(s)


if __name__ == "__main__":
    main()
