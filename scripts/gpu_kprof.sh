#!/usr/bin/env bash
# ncu --set full (with source) of one kernel of the val_real_probe run: usage gpu_kprof.sh TAG KERNEL_REGEX [SKIP]
set -u
TAG=${1:-r2k}; KRE=${2:-cnn_peaks_kernel}; SKIP=${3:-0}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o $OUT/${TAG}_prof -f python scripts/val_real_probe.py > $OUT/${TAG}_prof.log 2>&1
tail -2 $OUT/${TAG}_prof.log | cut -c1-200
