#!/usr/bin/env bash
# fingerprint kernel on the 4000 real reads: shipped library, comparison libraries under lib/var (old kernel, phase clock)
set -u
OUT=gpurun_out; TAG=${1:-fpr}
mkdir -p $OUT
: > $OUT/${TAG}_real.log
echo "== default" >> $OUT/${TAG}_real.log
FP_REAL=1 timeout 200 python scripts/fp_probe.py >> $OUT/${TAG}_real.log 2>&1
for f in warpdemux_b200/lib/var/libwdxfp_*.so; do
  echo "== $f" >> $OUT/${TAG}_real.log
  WDX_B200_LIB=$PWD/$f FP_REAL=1 timeout 200 python scripts/fp_probe.py >> $OUT/${TAG}_real.log 2>&1
done
grep -E "^==|^\{|rror" $OUT/${TAG}_real.log
