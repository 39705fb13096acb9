"""Other configs of BASELINE.json through the public API on one B200 (device time of a resident batch):
cfg1 (1,000 x 500 nt), cfg3 sample (long lncRNAs, 10-100 kb), cfg4 (2 kb sequences, W = 20 / 70 / 150),
plus a parity spot-check of one 20 kb sequence against the oracle's exact-math twin.
usage: sweep.py [out.json]"""
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


res = {}
res["cfg1_1000x500_W70"] = timed(workloads.cfg1(), 70)
res["cfg3_first64_W70"] = timed(workloads.cfg3(first=64), 70)
for W in (20, 70, 150):
    res[f"cfg4_2000x2kb_W{W}"] = timed(workloads.cfg4(first=2000), W, reps=2)
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
