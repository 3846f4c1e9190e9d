#!/usr/bin/env bash
# fingerprint stage: memcheck on the small parity test, full GPU suite, throughput probe
set -u
TAG=${1:-fp}
OUT=gpurun_out
mkdir -p $OUT
echo "== compute-sanitizer (golden fingerprint test)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fingerprint.py -x -q -k "golden or synthetic" > $OUT/${TAG}_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -15 $OUT/${TAG}_sanitizer.log
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 $OUT/${TAG}_pytest.log
echo "== fp probe"
timeout 300 python scripts/fp_probe.py 2>&1 | tee $OUT/${TAG}_probe.log
