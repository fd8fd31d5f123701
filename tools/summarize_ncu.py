"""Turn ncu exports (gpurun_out/) into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  profiles/r01_launches_R7.txt
    python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/r01_full_A_R6.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void smf::", "").replace("smf::", "")
        t = float(r[vi].replace(",", "")) / 1000.0
        a = agg[name]
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({len(rows) - 1} launches, {tot / 1000:.2f} ms; cold-cache, serialised)\n")
        f.write(f"# source: {src}\n{'kernel':45s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'max us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:45s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot * 100:6.1f}% {v[2]:10.1f}\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: {src}\n")
        for r in data:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"{k:90s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n")
            for i, k in enumerate(hdr):
                if "average_warps_issue_stalled" in k and "not_issued" not in k:
                    f.write(f"stall/issue {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {r[i]:>10s}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
