#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, probes, bench, ncu launch list and a full
# ncu capture of the dominant kernel.  Run as
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log

echo "== perf probe"
timeout 300 python scripts/perf_probe.py > $OUT/${TAG}_probe.log 2>&1
cat $OUT/${TAG}_probe.log

echo "== bench (N=1)"
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cat $OUT/${TAG}_bench_ref.json

echo "== ncu launch list (same bench command, smaller step count; times are cold-cache/serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-modes --no-extras \
    > $OUT/${TAG}_bench_under_ncu.log 2>&1
tail -2 $OUT/${TAG}_launches.csv

echo "== ncu --set full on the fused kernel (1 M reads so one replay pass stays short)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dtw_svc_kernel -s 2 -c 2 \
    -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --reads-per-gpu 1000000 --mode fast \
    --no-cpu-baseline --no-extra-modes --no-extras > $OUT/${TAG}_prof.log 2>&1
ls -la $OUT | tail -20
echo "== ncu --set full on the fingerprint kernel"
FP_REPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fingerprint_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_fpprof -f python scripts/fp_probe.py > $OUT/${TAG}_fpprof.log 2>&1
tail -2 $OUT/${TAG}_fpprof.log
