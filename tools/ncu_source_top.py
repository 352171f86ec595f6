"""Summarise `ncu --page source --csv` output: hottest SASS instructions by warp-stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "# Samples" in r)
data = [r for r in rows if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("sass rows", len(data), "total samples", tot, "warp instructions", sum(int(r[ie]) for r in data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    st = sorted(((hdr[i][6:], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% ie={r[ie]:>8s}  {r[src].strip()[:64]:64s} {st}")
