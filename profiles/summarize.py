"""Turns the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/<round>/.

  python profiles/summarize.py r1 gpurun_out/launches_r1.csv gpurun_out/prof_r1_full.ncu-rep|raw.csv <nt_per_launch>

 * launches csv  : `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of the bench command
 * .ncu-rep      : `ncu --set full --clock-control none --import-source on` of profiles/prof_driver.py
Outputs: launches_bench_raw.csv (copy), launches_bench_summary.csv (per-kernel shares), ncu_full_summary.json
(selected counters per kernel), traffic.json (DRAM bytes per launch of the dominant kernel; bench.py reads it).
"""
import csv
import json
import os
import shutil
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = name.replace("void <unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.split("(")[0]


def main():
    rnd, launches, rep, nt = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), rnd)
    os.makedirs(out, exist_ok=True)
    # ---- launch list -> shares
    shutil.copy(launches, os.path.join(out, "launches_bench_raw.csv"))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = {}
    for r in rows:
        if r is hdr or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].strip(), 1e-6)
        k = short(r[ik])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + v * scale)
    whole = sum(t for _, t in tot.values())
    with open(os.path.join(out, "launches_bench_summary.csv"), "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` "
                "(gpu__time_duration.sum, --clock-control none)\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{t:.3f},{100 * t / whole:.1f}\n")
    # ---- full capture -> selected counters
    # (a .csv argument is the `ncu -i <rep> --page raw --csv` output made on the GPU box: a report with sources can
    # exceed what gpurun copies back)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(
        ["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    summ = []
    for r in rr[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        e = {"Kernel Name": d["Kernel Name"]}
        for k in KEEP:
            if k in d:
                e[k] = f"{d[k]} {u[k]}".strip()
        summ.append(e)
    json.dump(summ, open(os.path.join(out, "ncu_full_summary.json"), "w"), indent=1)
    # ---- DRAM traffic of the dominant kernel (longest launch in the full capture)
    def ms(e):
        return float(e["gpu__time_duration.sum"].split()[0])
    def gb(e, k):
        v, unit = e[k].split()[:2]
        return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
    dom = max(summ, key=ms)
    json.dump({
        "source": f"profiles/{rnd}/ncu_full_summary.json (ncu --set full --clock-control none, "
                  "python profiles/prof_driver.py 1536 65536 = the bench step's batch)",
        "kernel": short(dom["Kernel Name"]),
        "dram_bytes_read_per_launch": int(gb(dom, "dram__bytes_read.sum")),
        "dram_bytes_write_per_launch": int(gb(dom, "dram__bytes_write.sum")),
        "nt_per_launch": nt,
    }, open(os.path.join(out, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(out, "launches_bench_summary.csv")).read())
    print(open(os.path.join(out, "traffic.json")).read())


if __name__ == "__main__":
    main()
