"""Dynamic instruction mix and stall samples per opcode from an ncu report captured with --import-source on.
  python profiles/sass_mix.py report.ncu-rep kernel_regex"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]
ci, cs, csm = cols.index("Instructions Executed"), cols.index("Source"), cols.index("# Samples")
inst, samp = defaultdict(int), defaultdict(int)
tot_i = tot_s = 0
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= ci or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    src = r[cs].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0].rstrip(";")
    n, s = int(r[ci]), int(r[csm])
    inst[op] += n
    samp[op] += s
    tot_i += n
    tot_s += s
    seq.append((n, s, src))
print(f"total warp instructions {tot_i}, samples {tot_s}")
for op, n in sorted(inst.items(), key=lambda kv: -kv[1])[:24]:
    print(f"  {op:10s} {n:12d} {100.0 * n / tot_i:5.1f}% inst   {100.0 * samp[op] / max(tot_s, 1):5.1f}% samples")
# regions: split the instruction stream into 40 equal chunks by address order, print inst/sample share of each
if len(sys.argv) > 3:
    k = int(sys.argv[3])
    step = (len(seq) + k - 1) // k
    for a in range(0, len(seq), step):
        part = seq[a:a + step]
        ni, ns = sum(p[0] for p in part), sum(p[1] for p in part)
        print(f"  [{a:5d}..{a + len(part):5d}) inst {100.0 * ni / tot_i:5.1f}%  samples {100.0 * ns / tot_s:5.1f}%   {part[0][2][:50]}")
