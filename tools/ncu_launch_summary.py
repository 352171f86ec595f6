"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last N launches = one step)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
names = [(r[ki], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1)) for r in data if len(r) > vi]
if last:
    names = names[-last:]
agg = collections.OrderedDict()
for n, t in names:
    short = re.sub(r"\(.*", "", n)
    short = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::|igemm::", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"launches {len(names)}, total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% x{v[0]:3d}  {k}")
