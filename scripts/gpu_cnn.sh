#!/usr/bin/env bash
# GPU visit for the boundary-CNN stage: parity tests (each under its own timeout: the tcgen05 kernel is new).
set -u
TAG=${1:-cnn}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_cnn.py -q -x -s -k "prepare or crafted or exact_mode or error" > $OUT/${TAG}_pytest_a.log 2>&1; echo "pytest A rc=$?"
tail -15 $OUT/${TAG}_pytest_a.log
timeout 300 python -m pytest tests/test_gpu_cnn.py -q -x -s -k "fast_mode or perturbed" > $OUT/${TAG}_pytest_b.log 2>&1; echo "pytest B rc=$?"
tail -25 $OUT/${TAG}_pytest_b.log
timeout 300 python scripts/cnn_probe.py > $OUT/${TAG}_probe.log 2>&1; echo "probe rc=$?"
tail -12 $OUT/${TAG}_probe.log
