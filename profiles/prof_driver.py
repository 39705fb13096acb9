"""Small fixed workload for ncu / quick experiments: cfg2 transcripts through the public API.
usage: prof_driver.py [n_seqs] [max_batch_MiB] [W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from priblast_b200 import Raccess, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
W = int(sys.argv[3]) if len(sys.argv) > 3 else 70
seqs = workloads.cfg2(first=n)
with Raccess(W, 5, max_batch_bytes=mb << 20) as r:
    nt = r.stage(seqs)
    r.compute()
    r.sync()
    c0 = r.counters()
    r.compute()
    r.sync()
    c = r.counters()
    ms = c["kernel_ms"] - c0["kernel_ms"]
    ph = {k: round(c["phase_ms"][k] - c0["phase_ms"][k], 2) for k in c["phase_ms"]}
    print(f"n={n} nt={nt} budget={mb}MiB batches={c['batches'] - c0['batches']} ms={ms:.2f} nt/s={nt / ms * 1e3:.3e} {ph}")
