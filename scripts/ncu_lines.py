"""Per-source-line totals (samples, warp instructions) of one kernel from
    ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > x.csv
    python scripts/ncu_lines.py x.csv [top]
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = defaultdict(lambda: [0, 0, ""])
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue  # keep only the per-line summary rows (address "-")
    try:
        s, i = int(r[i_s]), int(r[i_i])
    except ValueError:
        continue
    a = agg[(cur, int(r[0]))]
    a[0] += s
    a[1] += i
    a[2] = r[1].strip()[:90]
ts = sum(a[0] for a in agg.values())
ti = sum(a[1] for a in agg.values())
print(f"total samples {ts}  warp instructions {ti}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% ins  {f}:{ln}  {a[2]}")

if len(sys.argv) > 3:  # phase table: "name:lo-hi,lo-hi;name2:..." over fingerprint_kernel.cuh line ranges
    print()
    for spec in sys.argv[3].split(";"):
        name, rng = spec.split(":")
        s = i = 0
        for part in rng.split(","):
            lo, hi = map(int, part.split("-"))
            for (f, ln), a in agg.items():
                if f == "fingerprint_kernel.cuh" and lo <= ln <= hi:
                    s += a[0]
                    i += a[1]
        print(f"{name:14s} {s/ts*100:5.1f}% samples {i/ti*100:5.1f}% instructions")
    s = sum(a[0] for (f, ln), a in agg.items() if f != "fingerprint_kernel.cuh")
    i = sum(a[1] for (f, ln), a in agg.items() if f != "fingerprint_kernel.cuh")
    print(f"{'other files':14s} {s/ts*100:5.1f}% samples {i/ti*100:5.1f}% instructions")
