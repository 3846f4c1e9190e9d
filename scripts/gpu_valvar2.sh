#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/valvar2.log
for f in warpdemux_b200/lib/var/libwdx_*.so; do
  WDX_B200_LIB=$PWD/$f timeout 200 python scripts/val_real_probe.py 2>&1 | tail -1 >> $OUT/valvar2.log
  WDX_B200_LIB=$PWD/$f timeout 300 python -m pytest tests/test_validate.py tests/test_gpu_llr.py -m gpu -q 2>&1 | tail -1 >> $OUT/valvar2.log
done
cat $OUT/valvar2.log
