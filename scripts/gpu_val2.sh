#!/usr/bin/env bash
# validation-stage iteration: parity tests of everything that touches validate_kernel, then the real-read timing probe
set -u
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_validate.py tests/test_gpu_llr.py tests/test_real_reads.py tests/test_io_results.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 300 python scripts/val_real_probe.py 2>&1 | tail -1 | tee $OUT/${TAG}_val_probe.json
