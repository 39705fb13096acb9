"""Wall time of the whole `pRIblast_b200 db` front-end (FASTA in, five database files out) on cfg2 samples."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from priblast_b200 import workloads
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# usage: db_e2e.py [n_transcripts] [cfg2|cfg3] [extra front-end options ...]   (cfg3: -c 500 as in BASELINE.json)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
extra = sys.argv[3:] if len(sys.argv) > 3 else (["-c", "500"] if cfg == "cfg3" else [])
seqs = workloads.cfg3(first=n) if cfg == "cfg3" else workloads.cfg2(first=n)
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "in.fa")
    workloads.write_fasta(fa, seqs, prefix="t")
    nt = sum(len(s) for s in seqs)
    env = dict(os.environ, PRIB_DB_TIMING="1")
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(root, "priblast_b200", "pRIblast_b200"), "db", "-i", fa, "-o", os.path.join(d, "db"), *extra],
                       env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    print(r.stderr.strip())
    assert r.returncode == 0, r.stdout + r.stderr
    sizes = {e: os.path.getsize(os.path.join(d, "db." + e)) for e in ("acc", "seq", "ind", "nam", "bas")}
    print(f"{cfg} {' '.join(extra)}: {n} transcripts, {nt} nt: db wall {dt:.2f} s = {nt/dt:.3e} nt/s; files {sizes}")
