"""Last five library launches of an ncu --metrics gpu__time_duration.sum --csv log: (kernel, microseconds, grid)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        out.append((d["Kernel Name"][:40], float(d["Metric Value"]) / 1e3, d.get("Grid Size")))
for o in [x for x in out if "wdx" in x[0]][-int(sys.argv[2]) if len(sys.argv) > 2 else -5:]:
    print(o)
