#!/usr/bin/env bash
set -u
TAG=${1:-var}
shift
OUT=gpurun_out
mkdir -p $OUT
for v in "$@"; do
  echo "== variant $v"
  WDX_FAST_VARIANT=$v timeout 300 python scripts/variant_probe.py 2>&1 | tee -a $OUT/${TAG}_variants.log
done
echo "== parity under variant ${PARITY_VARIANT:-6}"
WDX_FAST_VARIANT=${PARITY_VARIANT:-6} timeout 600 python -m pytest tests/test_gpu_predict.py -x -q 2>&1 | tail -5
