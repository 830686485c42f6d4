"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, shares, first launches."""
import csv
import re
import sys
from collections import OrderedDict

src = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if r[col["Metric Name"]] == "gpu__time_duration.sum"]


def short(name):
    m = re.match(r"(?:void )?(?:celeste::)?(\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


agg = OrderedDict()
for r in data:
    k = short(r[col["Kernel Name"]])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[col["Metric Value"]]) / 1e3
tot = sum(a[1] for a in agg.values())
print("\n".join(sys.argv[2:]))
print()
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} launches={n:3d} total={us / 1e3:9.3f} ms avg={us / n:9.1f} us share_of_all_launches={us / tot:.3f}")
print("\nlaunch sequence (kernel, grid, us):")
seq = [r for r in data if "prep_image" not in r[col["Kernel Name"]]]
for r in seq[:12] + seq[-12:]:
    print(f"  {short(r[col['Kernel Name']]):34s} grid={r[col['Grid Size']]:>16s} {float(r[col['Metric Value']]) / 1e3:10.1f}")
