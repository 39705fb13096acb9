"""Times the exact engine (mode 2) on a cfg2 sample through the C ABI; prints nt/s and the phase split."""
import json
import sys
import time

sys.path.insert(0, ".")
from priblast_b200 import Raccess, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seqs = workloads.cfg2(first=n)
nt = sum(len(s) for s in seqs)
with Raccess(70, 5, mode=2) as r:
    r.run_batch(seqs[:8])  # warm-up
    c0 = r.counters()
    t0 = time.perf_counter()
    r.run_batch(seqs)
    wall = time.perf_counter() - t0
    c1 = r.counters()
ph = {k: round(c1["phase_ms"][k] - c0["phase_ms"][k], 2) for k in c1["phase_ms"]}
kms = c1["kernel_ms"] - c0["kernel_ms"]
print(json.dumps({"engine": "exact", "sequences": n, "nt": nt, "wall_s": round(wall, 3), "kernel_ms": round(kms, 2),
                  "nt_per_s_device": round(nt / (kms / 1e3)), "nt_per_s_e2e": round(nt / wall), "phase_ms": ph,
                  "launches": c1["kernel_launches"] - c0["kernel_launches"]}))
