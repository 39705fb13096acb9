#!/usr/bin/env python
"""TEST/BENCH INFRASTRUCTURE — algorithmic work per nucleotide measured with the oracle's
term counter on samples of the benchmark workloads; writes priblast_b200/data/work_per_nt.json, which
bench.py uses to turn nt/s into roofline operations/s.

term      = one summand entering a log-sum / probability sum of the reference recurrences
            (logsumexp calls + explicit expd calls, direct-path accounting of the biloop for every Z)
reduction = 14 (W-1) per nt (12 band variables + 2 outer-array rows) + the final logs
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle_py import OracleLib  # noqa: E402
from priblast_b200 import workloads  # noqa: E402


def measure(o, seqs, W, delta=5):
    tot = dict(nt=0, terms=0, loop_energy=0, cells=0, logs=0, inside=0, outside=0, access=0)
    for s in seqs:
        c = o.count_terms(s.decode(), W, delta)
        tot["nt"] += len(s)
        tot["terms"] += c["lse_inside"] + c["lse_outside"] + c["lse_access"] + c["expd_access"]
        tot["inside"] += c["lse_inside"]
        tot["outside"] += c["lse_outside"]
        tot["access"] += c["lse_access"] + c["expd_access"]
        tot["loop_energy"] += c["loop_energy"]
        tot["cells"] += c["cells"]
        tot["logs"] += c["final_logs"]
    nt = tot["nt"]
    red = 14 * (W - 1) + tot["logs"] / nt
    return dict(sample_nt=nt, sample_seqs=len(seqs), terms_per_nt=tot["terms"] / nt,
                reductions_per_nt=red, terms_inside_per_nt=tot["inside"] / nt, terms_outside_per_nt=tot["outside"] / nt,
                terms_access_per_nt=tot["access"] / nt, loop_energy_calls_per_nt=tot["loop_energy"] / nt,
                cells_per_nt=tot["cells"] / nt,
                sfu_ops_per_nt=tot["terms"] / nt + red, fp32_instr_per_nt=6 * tot["terms"] / nt)


def main():
    o = OracleLib()
    out = {}
    out["cfg2_W70"] = measure(o, workloads.cfg2(first=20), 70)
    out["cfg1_W70"] = measure(o, workloads.cfg1()[:40], 70)
    for W in (20, 70, 150):
        out[f"cfg4_W{W}"] = measure(o, workloads.cfg4(first=6 if W < 150 else 3), W)
    out["cfg3_W70"] = measure(o, [workloads.cfg3(first=1)[0][:12000]], 70)
    path = os.path.join(HERE, "..", "priblast_b200", "data", "work_per_nt.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
