"""Small fixed workload for ncu: one device batch of cfg2 transcripts through the public API."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from priblast_b200 import Raccess, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seqs = workloads.cfg2(first=n)
with Raccess(70, 5, max_batch_bytes=16 << 30) as r:
    r.stage(seqs)
    r.compute()
    r.sync()
    c = r.counters()
    print(n, "seqs", c["nucleotides"], "nt", c["kernel_ms"], "ms", c["phase_ms"])
