"""Other configs of BASELINE.json through the public API on one B200 (device time, inputs resident), AT THEIR FULL
SIZES: cfg1 (1,000 x 500 nt), cfg3 (2,000 lncRNAs of 10-100 kb, 7.9e7 nt), cfg4 (10,000 x 2 kb, W = 20 / 70 / 150),
plus a parity spot-check of one 20 kb sequence against the oracle's exact-math twin.
usage: sweep.py [out.json] [quick]   (quick: cfg3 first 64, cfg4 first 2,000 — the round-1 sample sizes)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
from priblast_b200 import Raccess, workloads


def timed(seqs, W, delta=5, reps=3, **kw):
    with Raccess(W, delta, **kw) as r:
        nt = r.stage(seqs)
        r.compute(); r.sync()
        best, rerun, batches, phases = None, 0, 0, None
        for _ in range(reps):  # per-repetition device time; the best one is reported
            c0 = r.counters()
            r.compute(); r.sync()
            c1 = r.counters()
            ms = c1["kernel_ms"] - c0["kernel_ms"]
            if best is None or ms < best:
                best = ms
                rerun = int(c1["fp64_rerun_sequences"] - c0["fp64_rerun_sequences"])
                batches = int(c1["batches"] - c0["batches"])
                phases = {k: round(c1["phase_ms"][k] - c0["phase_ms"][k], 3) for k in c1["phase_ms"]}
        return {"sequences": len(seqs), "nt": nt, "W": W, "ms": round(best, 3), "nt_per_s": nt / best * 1e3,
                "batches": batches, "fp64_rerun": rerun, "phase_ms_fp32_pass": phases}


quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
res = {}
res["cfg1_1000x500_W70"] = timed(workloads.cfg1(), 70)
res["cfg3_first64_W70"] = timed(workloads.cfg3(first=64), 70)
if not quick:
    res["cfg3_2000_W70"] = timed(workloads.cfg3(), 70, reps=2)
n4 = 2000 if quick else 10000
for W in (20, 70, 150):
    res[f"cfg4_{n4}x2kb_W{W}"] = timed(workloads.cfg4(first=n4), W, reps=2)
wk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "priblast_b200", "data",
                                 "work_per_nt.json")))
for key, w in (("cfg4_%dx2kb_W20" % n4, "cfg4_W20"), ("cfg4_%dx2kb_W70" % n4, "cfg4_W70"), ("cfg4_%dx2kb_W150" % n4, "cfg4_W150")):
    if w in wk:  # fraction of the SURVEY 8(d) SFU yardstick (4.64e12 MUFU/s nominal)
        res[key]["sfu_frac_nominal"] = wk[w]["sfu_ops_per_nt"] * res[key]["nt_per_s"] / 4.653e12
# parity spot check on a long sequence (log-path regime, Z ~ 5,000)
from oracle_py import OracleLib
seq = workloads.cfg3(first=1)[0][:20000]
with Raccess(70, 5) as r:
    a, c = r.run(seq)
t0 = time.perf_counter()
ea, ec = OracleLib().run_exact(seq, 70, 5)
res["parity_20kb_vs_exact_twin"] = {"max_abs_acc": float(np.abs(a - ea).max()), "max_abs_cond": float(np.abs(c - ec).max()),
                                    "oracle_seconds": round(time.perf_counter() - t0, 1)}
out = json.dumps(res, indent=1)
print(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(out)
