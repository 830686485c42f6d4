"""profiles/ncu_traffic.json from `ncu --set full` captures of the bench workload (tools/profile_step.py 10 <mode> 3):
per-launch DRAM traffic, FP64 pipe / issue activity and the EXECUTED flop count (2 x DFMA + DMUL + DADD thread
instructions) of the dominant kernel of each mode.  bench.py reads it for roofline.traffic / roofline.executed.

usage: python tools/ncu_traffic.py <key>=<file.ncu-rep> ...   e.g.  "march_kernel<1>"=gpurun_out/prof_march.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
import re
jobs = []
for arg in sys.argv[1:]:
    key, f = arg.split("=", 1) if "=" in arg else ("auto", arg)
    raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    for vals in rows[2:]:
        jobs.append((key, rows[0], rows[1], vals))
for key, hdr, units, vals in jobs:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    if key == "auto":                  # "void unit_walk_kernel<2>(PlanDev, ...)" -> "unit_walk_kernel<2>"
        m = re.search(r"(\w+(?:<[^>(]*>)?)\s*\(", d["Kernel Name"])
        key = m.group(1)

    def val(k):
        return float(d[k])

    def to_bytes(k):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
        return val(k) * scale

    dur_us = val("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "s": 1e6, "ns": 1e-3}[u["gpu__time_duration.sum"]]
    cyc = val("sm__cycles_elapsed.avg")
    dfma = val("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
    dmul = val("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
    dadd = val("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
    flop = (2 * dfma + dmul + dadd) * cyc
    mode = key[-2] if key.endswith(">") else "2"
    res[key] = {
        "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
        "sources": int(val("launch__grid_size")) if ("march" in key or "task" in key) else 10000,
        "launch": f"the bench workload: one launch over the 10 000-source stripe (tools/profile_step.py 10 {mode} 3), "
                  "ncu --set full --clock-control none",
        "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": val("launch__registers_per_thread"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "duration_us": dur_us,
        "executed_flop_per_launch": flop,
        "executed_tflops": flop / (dur_us * 1e-6) / 1e12,
        "sass_thread_inst_per_cycle": {"dfma": dfma, "dmul": dmul, "dadd": dadd},
        "sm_cycles_elapsed": cyc,
    }
    print(key, json.dumps(res[key], indent=1))
json.dump(res, open(out_path, "w"), indent=1)
