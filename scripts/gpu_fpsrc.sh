#!/usr/bin/env bash
# ncu --set full capture (with source correlation) of the fingerprint kernel; read here with
#   ncu -i gpurun_out/<tag>_fpprof.ncu-rep --page source --csv
set -u
TAG=${1:-fpsrc}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/fp_probe.py 2>&1 | tee $OUT/${TAG}_probe.log
FP_REPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fingerprint_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_fpprof -f python scripts/fp_probe.py > $OUT/${TAG}_fpprof.log 2>&1
tail -3 $OUT/${TAG}_fpprof.log
ls -la $OUT
