#!/usr/bin/env python
"""Turn gpurun_out/<tag>_launches.csv + <tag>_fused.ncu-rep into tracked summaries under profiles/.

    tools/make_profile_summary.py <tag> <kernel-substring> [mangled-substring-for-the-line-tool]
"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, kernel = sys.argv[1], sys.argv[2]
line_kernel = sys.argv[3] if len(sys.argv) > 3 else kernel  # mangled-name substring that selects ONE instantiation
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1] / 1e3:.1f},{100 * v[1] / tot:.2f},{v[1] / 1e3 / v[0]:.1f}\n")
subprocess.run(["cp", os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv"), os.path.join(out_dir, f"{tag}_launches.csv")])

# ---- raw metrics of the top kernel ---------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_fused.ncu-rep")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines()))
names, units, vals = r[0], r[1], r[-1]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")
with open(os.path.join(out_dir, f"{tag}_fused_metrics.csv"), "w") as f:
    f.write("metric,unit,value\n")
    for n, u, v in zip(names, units, vals):
        if n in want or n.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in n:
            f.write(f"{n},{u},{v}\n")

# ---- per-source-line attribution -------------------------------------------------------------------
lib = os.path.join(ROOT, "cumind_b200", "_lib", "libcumind_b200.so")
res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, lib, line_kernel, "40"], capture_output=True, text=True)
open(os.path.join(out_dir, f"{tag}_fused_by_line.txt"), "w").write(res.stdout)
print(open(os.path.join(out_dir, f"{tag}_launches_summary.csv")).read())
print(open(os.path.join(out_dir, f"{tag}_fused_metrics.csv")).read())
