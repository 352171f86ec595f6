"""Per-kernel summary of ONE step from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list (one CSV row per metric and launch).  usage: ncu_step_summary.py <csv> <launches per step> [marker]
The last <launches per step> launches are summarised; with [marker] (a kernel-name substring that ends a step, e.g.
adam_kernel) the last complete step between two markers is used instead."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
per_step = int(sys.argv[2])
marker = sys.argv[3] if len(sys.argv) > 3 else None
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ii, ki, mi, ui, vi = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    launches.setdefault(r[ii], {"name": r[ki]})[r[mi]] = v * scale
ls = list(launches.values())
if marker:
    idx = [i for i, l in enumerate(ls) if marker in l["name"]]
    ls = ls[idx[-2] + 1: idx[-1] + 1]
else:
    ls = ls[-per_step:]
agg = collections.OrderedDict()
for l in ls:
    short = re.sub(r"\(.*", "", l["name"])
    short = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::|igemm::", "", short)[:100]
    a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += l.get("gpu__time_duration.sum", 0.0)
    a[2] += l.get("dram__bytes_read.sum", 0.0)
    a[3] += l.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(ls)}, total {tot:.1f} us (serialised, under ncu)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{a[1]:10.1f} us {100 * a[1] / tot:5.1f}% x{a[0]:3d}  DRAM {a[2]:9.1f} MB read {a[3]:9.1f} MB written  {k}")
ig = [a for k, a in agg.items() if k.startswith("igemm_kernel")]
if ig:
    n = sum(a[0] for a in ig)
    rd, wr = sum(a[2] for a in ig), sum(a[3] for a in ig)
    print(f"igemm family: {n} launches, {rd / 1e3:.2f} GB read + {wr / 1e3:.2f} GB written per step = {(rd + wr) / n:.1f} MB per launch")
