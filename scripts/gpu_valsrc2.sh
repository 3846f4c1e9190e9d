#!/usr/bin/env bash
# ncu --set full capture (with source correlation) of validate_kernel and llr_kernel on the REAL reads
set -u
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 2 -c 1 \
    -o $OUT/${TAG}_valprof -f python scripts/val_real_probe.py > $OUT/${TAG}_valprof.log 2>&1
tail -2 $OUT/${TAG}_valprof.log | cut -c1-300
# launches of the third configuration (verdict_only + LLR): validate, llr, validate, llr, validate per call; 5 calls
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llr_kernel -s 3 -c 2 \
    -o $OUT/${TAG}_llrprof -f python scripts/val_real_probe.py > $OUT/${TAG}_llrprof.log 2>&1
tail -2 $OUT/${TAG}_llrprof.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_val_launches.csv python scripts/val_real_probe.py > /dev/null 2>&1
ls -la $OUT | tail -5
