#!/usr/bin/env bash
# Round 2, visit c: bench (both arms, all extras), launch list, ncu --set full of the benchmarked fused launch.
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g >> $OUT/${TAG}_gpu.txt
( time timeout 1200 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2> $OUT/${TAG}_bench.time
cut -c1-1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.time
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err ) 2> $OUT/${TAG}_bench_ref.time
cut -c1-700 $OUT/${TAG}_bench_ref.json; cat $OUT/${TAG}_bench_ref.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-extra-modes --e2e-steps 1 --no-chain-stream > $OUT/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dtw_svc_kernel --launch-skip 18 --launch-count 1 -f -o $OUT/${TAG}_dtw_svc \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-extra-modes --e2e-steps 1 --no-chain-stream > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-300
ls -la $OUT | tail -8
