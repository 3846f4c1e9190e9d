#!/usr/bin/env bash
# GPU visit for the validation step: its parity tests, the chained real-read test, the probe, then the whole GPU suite.
set -u
TAG=${1:-r01l}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_validate.py tests/test_real_reads.py -m gpu -x -q > $OUT/${TAG}_pytest_val.log 2>&1; echo "pytest val rc=$?" | tee -a $OUT/${TAG}_pytest_val.log
tail -30 $OUT/${TAG}_pytest_val.log
timeout 600 python scripts/validate_probe.py > $OUT/${TAG}_validate_probe.jsonl 2> $OUT/${TAG}_validate_probe.err; echo "probe rc=$?"
cat $OUT/${TAG}_validate_probe.jsonl; tail -5 $OUT/${TAG}_validate_probe.err
if [ "${2:-full}" = "full" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
fi
