"""Where the end-to-end time of one prib_acc_run-equivalent goes: marshal / stage / compute / fetch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from priblast_b200 import Raccess, workloads, packed_layout

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
seqs = workloads.cfg2(first=n)
lens = np.array([len(s) for s in seqs])
_, _, total = packed_layout(lens)
out = torch.empty(total, dtype=torch.float32).pin_memory().numpy()
with Raccess(70, 5) as r:
    for it in range(4):
        t0 = time.perf_counter(); r.stage(seqs); t1 = time.perf_counter()
        r.compute(); r.sync(); t2 = time.perf_counter()
        r.fetch(out); t3 = time.perf_counter()
        r.run_batch(seqs, out=out); t4 = time.perf_counter()
        print(f"stage {1e3*(t1-t0):.2f} ms  compute {1e3*(t2-t1):.2f}  fetch {1e3*(t3-t2):.2f}  | run_batch {1e3*(t4-t3):.2f}")
    c = r.counters()
    print({k: c[k] for k in ("h2d_ms", "d2h_ms", "kernel_ms", "h2d_bytes", "d2h_bytes")})
