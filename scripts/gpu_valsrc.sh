#!/usr/bin/env bash
# ncu --set full capture (with source correlation) of validate_kernel; read here with
#   ncu -i gpurun_out/<tag>_valprof.ncu-rep --page source --csv --print-source cuda,sass
set -u
TAG=${1:-valsrc}
OUT=gpurun_out
mkdir -p $OUT
VAL_REPS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_valprof -f python scripts/validate_probe.py > $OUT/${TAG}_valprof.log 2>&1
tail -3 $OUT/${TAG}_valprof.log
ls -la $OUT | tail -5
