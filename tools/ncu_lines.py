"""Per CUDA source line: stall samples and executed warp instructions of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys
from collections import defaultdict

f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
import os
extra = os.environ.get("NCU_ARGS", "").split()        # e.g. NCU_ARGS="-k regex:unit_walk"
out = subprocess.run(["ncu", "-i", f] + extra + ["--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, hdr, rows = None, None, []
agg = defaultdict(lambda: [0, 0, 0, ""])   # samples, inst, thread inst, text
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    key = (cur_file, int(r[0]))
    a = agg[key]
    num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
    a[0] += num(r[hdr["# Samples"]])
    a[1] += num(r[hdr["Instructions Executed"]])
    a[2] += num(r[hdr["Thread Instructions Executed"]])
    a[3] = r[1].strip()[:90]
tot_s = sum(a[0] for a in agg.values())
tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}  warp instructions {tot_i}")
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lanes = a[2] / a[1] if a[1] else 0
    print(f"{fn:22s}:{ln:4d} samples {100*a[0]/tot_s:5.2f}%  inst {100*a[1]/tot_i:5.2f}%  lanes {lanes:4.1f}  {a[3]}")
