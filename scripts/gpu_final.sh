#!/usr/bin/env bash
# Round-end style visit: GPU tests, smoke, bench (both arms), ncu launch list of the bench command.
set -u
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
bash scripts/gpu_quick.sh $TAG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-extra-modes --e2e-steps 1 > $OUT/${TAG}_bench_under_ncu.log 2>&1
tail -2 $OUT/${TAG}_bench_under_ncu.log | cut -c1-300
timeout 300 python scripts/latency_probe.py 2>&1 | grep numpy_api | tee $OUT/${TAG}_latency.jsonl
