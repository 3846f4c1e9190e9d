#!/usr/bin/env bash
# Round 2, visit a: LLR fallback tests first (verbose), then the whole GPU suite.
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_llr.py -m gpu -q -x > $OUT/${TAG}_pytest_llr.log 2>&1; echo "llr rc=$?" | tee -a $OUT/${TAG}_pytest_llr.log
tail -40 $OUT/${TAG}_pytest_llr.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_llr.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
