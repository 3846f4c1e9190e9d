#!/usr/bin/env bash
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_validate.py tests/test_gpu_llr.py tests/test_real_reads.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python scripts/val_real_probe.py 2>&1 | tail -1 | tee $OUT/${TAG}_val_probe.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_val_launches.csv python scripts/val_real_probe.py > /dev/null 2>&1
python scripts/launch_tail.py $OUT/${TAG}_val_launches.csv
