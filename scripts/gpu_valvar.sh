#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/valvar.log
for f in warpdemux_b200/lib/var/libwdx_*.so; do
  for stride in 11500 16000; do
    echo "== $f stride $stride" >> $OUT/valvar.log
    WDX_B200_LIB=$PWD/$f VAL_ONLY=1 VAL_STRIDE=$stride timeout 120 python scripts/validate_probe.py >> $OUT/valvar.log 2>&1
  done
  WDX_B200_LIB=$PWD/$f timeout 200 python -m pytest tests/test_validate.py -m gpu -q 2>&1 | tail -1 >> $OUT/valvar.log
done
cat $OUT/valvar.log
