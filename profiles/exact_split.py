"""Exact engine: phase times on a direct-path set (Z <= 690) and on a log-sum-path set (Z > 690)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from priblast_b200 import Raccess

rng = np.random.default_rng(11)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
sets = {"direct_1500nt_x256": (1500, 256), "logsum_3500nt_x110": (3500, 110)}
with Raccess(70, 5, mode=2) as r:
    r.run_batch(["ACGUACGUGGCCAAUU" * 20])
    for name, (L, n) in sets.items():
        if which not in ("both", name.split("_")[0]):
            continue
        seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for _ in range(n)]
        c0 = r.counters()
        r.run_batch(seqs)
        c1 = r.counters()
        ph = {k: round(c1["phase_ms"][k] - c0["phase_ms"][k], 2) for k in c1["phase_ms"]}
        print(json.dumps({"set": name, "nt": L * n, "kernel_ms": round(c1["kernel_ms"] - c0["kernel_ms"], 1), "phase_ms": ph}))
