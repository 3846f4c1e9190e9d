#!/usr/bin/env bash
set -u
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
( time timeout 1200 python bench.py --steps 3 --warmup 3 --cadence-seconds 5 --single-call-reads 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2> $OUT/${TAG}_bench.time
cut -c1-600 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.time
