"""Times library variants (profiles/build_variants.py) on the bench step: 1,536 cfg2 transcripts, device-resident.

  python profiles/variants_time.py [name ...]        (default: every variant found; 'default' = the product library)
Each variant runs in its own process (PRIB_ACC_LIB selects the library)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
from priblast_b200 import Raccess, workloads
n, W = int(sys.argv[1]), int(sys.argv[2])
seqs = workloads.cfg2(first=n) if W == 70 else workloads.cfg4(first=n)
with Raccess(W, 5) as r:
    nt = r.stage(seqs)
    for _ in range(3):
        r.compute()
    r.sync()
    c0 = r.counters()
    reps = 5
    for _ in range(reps):
        r.compute()
    r.sync()
    c = r.counters()
    ms = (c["kernel_ms"] - c0["kernel_ms"]) / reps
    ph = {k: round((c["phase_ms"][k] - c0["phase_ms"][k]) / reps, 3) for k in c["phase_ms"]}
    out = r.fetch()
    chk = float(sum(np.float64(a).sum() + np.float64(b).sum() for a, b in out[:64]))
    print(json.dumps({"nt": nt, "ms": round(ms, 3), "nt_per_s": nt / ms * 1e3, "phases": ph, "checksum64": chk,
                      "rerun": c["fp64_rerun_sequences"] - c0["fp64_rerun_sequences"]}))
''' % ROOT


def main():
    names = sys.argv[1:]
    vdir = os.path.join(ROOT, "priblast_b200", "variants")
    if not names:
        names = ["default"] + sorted(os.path.basename(p)[len("libpriblast_acc_"):-3]
                                     for p in glob.glob(os.path.join(vdir, "libpriblast_acc_*.so")))
    n = int(os.environ.get("VT_NSEQ", "1536"))
    W = int(os.environ.get("VT_W", "70"))
    for name in names:
        env = dict(os.environ)
        if name != "default":
            env["PRIB_ACC_LIB"] = os.path.join(vdir, f"libpriblast_acc_{name}.so")
        p = subprocess.run([sys.executable, "-c", CHILD, str(n), str(W)], env=env, capture_output=True, text=True)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "FAILED: " + p.stderr[-400:]
        print(f"{name:>12}: {line}", flush=True)


if __name__ == "__main__":
    main()
