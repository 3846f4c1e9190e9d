#!/usr/bin/env bash
# launch list (ncu gpu__time_duration) of the whole chain probe; prints the per-kernel totals of the LAST device-resident 8000-read pass
set -u
TAG=${1:-r2cl}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_chain_launches.csv python scripts/chain_probe.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/${TAG}_chain_launches.csv")))
hdr=None; L=[]
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); L.append((d["Kernel Name"][:60], float(d["Metric Value"].replace(",",""))/1e3, d.get("Grid Size")))
# the last launch of the big fused dtw kernel marks the last 8000-read pass of the stage timing; walk back to the cnn_prepare before it
idx=[i for i,x in enumerate(L) if "validate_kernel" in x[0] and x[2].startswith("(296")]
end=len(L)
# print the last 60 launches
for x in L[-70:]: print("%-62s %9.1f us %s"%x)
PY
