#!/usr/bin/env bash
# ncu --set full capture (with source correlation) of the first big validate_kernel launch on the REAL reads
set -u
TAG=${1:-r2w}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 2 -c 1 \
    -o $OUT/${TAG}_valprof -f python scripts/val_real_probe.py > $OUT/${TAG}_valprof.log 2>&1
tail -2 $OUT/${TAG}_valprof.log | cut -c1-300
