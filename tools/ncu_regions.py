"""Stall samples of an .ncu-rep grouped by source-line RANGES (regions of a kernel) and by stall reason.
    python tools/ncu_regions.py rep file.cuh name:lo-hi [name:lo-hi ...]      (lines outside every range -> 'other')"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, fname = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    nm, rg = a.split(":")
    lo, hi = rg.split("-")
    regions.append((nm, int(lo), int(hi)))
import os
extra = os.environ.get("NCU_ARGS", "").split()        # e.g. NCU_ARGS="-k regex:unit_walk"
out = subprocess.run(["ncu", "-i", rep] + extra + ["--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
agg = defaultdict(lambda: defaultdict(int))
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    ln = int(r[0])
    reg = cur
    if cur == fname:
        reg = "other"
        for nm, lo, hi in regions:
            if lo <= ln <= hi:
                reg = nm
                break
    for i, h in enumerate(hdr):
        if h == "# Samples" or h == "Instructions Executed" or (h.startswith("stall_") and "Not Issued" not in h):
            v = r[i]
            if v and v.lstrip("-").isdigit():
                agg[reg][h] += int(v)
tot = sum(a["# Samples"] for a in agg.values())
toti = sum(a["Instructions Executed"] for a in agg.values())
keys = ["stall_no_inst", "stall_long_sb", "stall_wait", "stall_math", "stall_short_sb", "stall_mio", "stall_lg", "stall_not_selected",
        "stall_selected", "stall_branch_resolving", "stall_dispatch", "stall_barrier"]
print(f"{'region':24s} {'samples%':>8s} {'inst%':>6s}  " + " ".join(f"{k[6:][:9]:>9s}" for k in keys))
for reg, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    if a["# Samples"] < 0.002 * tot:
        continue
    print(f"{reg:24s} {100*a['# Samples']/tot:8.2f} {100*a['Instructions Executed']/toti:6.2f}  " +
          " ".join(f"{100*a[k]/tot:9.2f}" for k in keys))
