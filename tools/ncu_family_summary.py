"""Kernel families of an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, share, launches."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
cnt, tot = collections.Counter(), collections.Counter()
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")[:70]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    cnt[name] += 1
    tot[name] += v
T = sum(tot.values())
print(f"launches {sum(cnt.values())}, total {T / 1e3:.1f} us")
for n, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{v / 1e3:10.1f} us {100 * v / T:5.1f}% x{cnt[n]:4d}  {n}")
