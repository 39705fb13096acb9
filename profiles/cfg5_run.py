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
        ref_by_key = {a[:3]: a for a in rr}
        for name in ("fast", "exact"):
            gr, gt = out[name]
            got_by_key = {b[:3]: b for b in gr}
            common = [k for k in ref_by_key if k in got_by_key]
            only_ref = [ref_by_key[k] for k in ref_by_key if k not in got_by_key]
            only_got = [got_by_key[k] for k in got_by_key if k not in ref_by_key]
            res[f"hits_{name}"] = len(gr)
            res[f"structural_hit_list_identical_{name}"] = not only_ref and not only_got
            res[f"hits_common_{name}"] = len(common)
            res[f"hits_only_in_reference_{name}"] = len(only_ref)
            res[f"hits_only_in_gpu_db_{name}"] = len(only_got)
            # the hits that exist on one side only: their interaction energy against the `-e` cut-off of ris (-6 by
            # default, main.cpp): a hit whose energy sits within the .acc tolerance of the cut-off can fall either way
            res[f"one_sided_hits_interaction_energy_{name}"] = sorted(round(h[5], 4) for h in only_ref + only_got)[:20]
            if common:
                res[f"max_energy_difference_common_hits_{name}"] = max(
                    max(abs(ref_by_key[k][3] - got_by_key[k][3]), abs(ref_by_key[k][4] - got_by_key[k][4]),
                        abs(ref_by_key[k][5] - got_by_key[k][5])) for k in common)
            res[f"text_identical_{name}"] = rt == gt
            res[f"printed_lines_differing_{name}"] = len(set(rt) ^ set(gt)) // 2
    s = json.dumps(res, indent=1)
    print(s)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(s)


if __name__ == "__main__":
    main()
