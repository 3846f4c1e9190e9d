#!/usr/bin/env bash
# one iteration on the fingerprint kernel: parity tests that reach it, throughput probe, ncu --set full with source
set -u
TAG=${1:-fpi}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fingerprint.py tests/test_gpu_trna.py tests/test_real_reads.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/${TAG}_pytest.log
timeout 300 python scripts/fp_probe.py 2>&1 | tee $OUT/${TAG}_probe.log
if [ "${2:-prof}" = "prof" ]; then
FP_REPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fingerprint_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_fpprof -f python scripts/fp_probe.py > $OUT/${TAG}_fpprof.log 2>&1
tail -2 $OUT/${TAG}_fpprof.log
fi
