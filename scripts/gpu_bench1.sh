#!/usr/bin/env bash
# default single-GPU bench, both arms, as the driver runs them (+ the launch list of a short run for profiles/)
set -u
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_ref.err ) 2> $OUT/${TAG}_bench_ref.time
cut -c1-400 $OUT/${TAG}_bench_reference_arm.json; cat $OUT/${TAG}_bench_ref.time
( time timeout 1500 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench.err ) 2> $OUT/${TAG}_bench.time
cut -c1-700 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
tail -5 $OUT/${TAG}_launches_bench.csv | cut -c1-200
