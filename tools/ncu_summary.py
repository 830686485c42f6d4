"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]

for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print("==", f, d.get("Kernel Name", ("?",))[0][:70])
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k][0]:>16s} {d[k][1]}")
        for h in hdr:
            if "issue_stalled" in h and "per_issue_active" in h and d[h][0]:
                v = float(d[h][0])
                if v > 0.08:
                    print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
