#!/usr/bin/env bash
set -u
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_real_reads.py tests/test_gpu_llr.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/stream_probe.py 2>&1 | tail -1 | tee $OUT/${TAG}_stream.json
