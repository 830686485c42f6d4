"""Top stalled SASS instructions of an .ncu-rep source page: which instructions eat the samples, and why."""
import csv
import subprocess
import sys

f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
srt = sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
for r in srt:
    st = {s[6:]: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{r[col['Address']][-5:]} {int(r[col['# Samples']]):6d} {100*int(r[col['# Samples']])/tot:5.2f}% {r[col['Source']][:70]:70s} {st}")
