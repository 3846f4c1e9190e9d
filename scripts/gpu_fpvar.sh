#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/fpvar.log
for short in 1200 2900; do
  cap=$(( (short + 200 + 63) / 64 * 64 ))
  for f in warpdemux_b200/lib/var/libwdxfp_*.so; do
    echo "== $f short $short cap $cap" >> $OUT/fpvar.log
    WDX_B200_LIB=$PWD/$f FP_SHORT=$short FP_MAX_SLICE=$cap timeout 120 python scripts/fp_probe.py >> $OUT/fpvar.log 2>&1
  done
done
cat $OUT/fpvar.log
