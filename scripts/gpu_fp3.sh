#!/usr/bin/env bash
set -u
TAG=${1:-fp}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/${TAG}_pytest.log
echo "== fp probe"
timeout 300 python scripts/fp_probe.py 2>&1 | tee $OUT/${TAG}_probe.log
echo "== exact variants"
for v in 0 1; do WDX_EXACT_VARIANT=$v timeout 300 python scripts/perf_probe.py 2>&1 | grep exact | tee -a $OUT/${TAG}_exact.log; done
echo "== ncu fingerprint"
FP_REPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fingerprint_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_fpprof -f python scripts/fp_probe.py > $OUT/${TAG}_fpprof.log 2>&1
tail -3 $OUT/${TAG}_fpprof.log
