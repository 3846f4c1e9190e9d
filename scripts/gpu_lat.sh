#!/usr/bin/env bash
set -u
TAG=${1:-lat}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_predict.py -m gpu -x -q 2>&1 | tail -3
echo "--- small-batch path ON" | tee $OUT/${TAG}_latency.jsonl
timeout 300 python scripts/latency_probe.py 2>&1 | tee -a $OUT/${TAG}_latency.jsonl
echo "--- small-batch path OFF (WDX_NO_SMALL_PATH=1)" | tee -a $OUT/${TAG}_latency.jsonl
WDX_NO_SMALL_PATH=1 timeout 300 python scripts/latency_probe.py 2>&1 | grep numpy_api | tee -a $OUT/${TAG}_latency.jsonl
