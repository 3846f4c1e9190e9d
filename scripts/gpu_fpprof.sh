#!/usr/bin/env bash
set -u
OUT=gpurun_out; TAG=${1:-fpp}
mkdir -p $OUT
: > $OUT/${TAG}_phases.log
for short in 0 2900 1200; do
  cap=7040; [ $short -gt 0 ] && cap=$(( (short + 200 + 63) / 64 * 64 ))
  echo "== short $short cap $cap" >> $OUT/${TAG}_phases.log
  WDX_B200_LIB=$PWD/warpdemux_b200/lib/var/libwdxfp_prof.so FP_SHORT=$short FP_MAX_SLICE=$cap timeout 120 python scripts/fp_probe.py >> $OUT/${TAG}_phases.log 2>&1
done
cat $OUT/${TAG}_phases.log
