"""ncu CSV (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum``) of tools/profile_step.py ->
per-kernel-family time share AND measured DRAM traffic of one step; writes profiles/igemm_traffic.json (what bench.py prints
as roofline.traffic: the average over exactly the igemm launches of one bench step, same batch).

usage: python tools/ncu_traffic.py launches.csv <launches per step> [out.json] ["command string"]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
per_step = int(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
idc, ki, mi, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d = launch.setdefault(int(r[idc]), {"name": r[ki]})
    d[r[mi]] = v * scale
ids = sorted(launch)[-per_step:]
fam = collections.OrderedDict()
for i in ids:
    d = launch[i]
    n = re.sub(r"\(.*", "", d["name"])
    n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::|igemm::", "", n)
    f = fam.setdefault(n, {"launches": 0, "us": 0.0, "read": 0.0, "write": 0.0})
    f["launches"] += 1
    f["us"] += d.get("gpu__time_duration.sum", 0.0)
    f["read"] += d.get("dram__bytes_read.sum", 0.0)
    f["write"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(f["us"] for f in fam.values())
print(f"launches {len(ids)}, total {tot:.1f} us (serialised, under ncu)")
for n, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{f['us']:10.1f} us {100 * f['us'] / tot:5.1f}% x{f['launches']:3d}  DRAM {f['read'] / 1e6:9.1f} MB read {f['write'] / 1e6:9.1f} MB written  {n}")
ig = {n: f for n, f in fam.items() if n.startswith("igemm_kernel")}
nl = sum(f["launches"] for f in ig.values())
rd, wr = sum(f["read"] for f in ig.values()), sum(f["write"] for f in ig.values())
print(f"igemm family: {nl} launches, {rd / 1e9:.2f} GB read + {wr / 1e9:.2f} GB written per step = {(rd + wr) / nl / 1e6:.1f} MB per launch")
if len(sys.argv) > 3:
    out = {"dram_bytes_per_launch": (rd + wr) / nl, "launches_per_step": nl, "dram_read_bytes_per_step": rd,
           "dram_write_bytes_per_step": wr, "share_of_step_under_ncu": sum(f["us"] for f in ig.values()) / tot,
           "per_family": {n: {"launches": f["launches"], "dram_read_mb": f["read"] / 1e6, "dram_write_mb": f["write"] / 1e6,
                              "us": f["us"]} for n, f in ig.items()},
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                     + (sys.argv[4] if len(sys.argv) > 4 else "tools/profile_step.py 64 2")
                     + ": average over the igemm launches of ONE encode step at the bench batch (64)"}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
