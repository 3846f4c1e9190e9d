#!/usr/bin/env bash
# GPU visit r01f: wavefront DTW parity + probe, fingerprint parity after the shared-window t-test + probe.
set -u
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_predict.py -q -x -k "wavefront or generic" > $OUT/${TAG}_pytest_wf.log 2>&1; echo "pytest wavefront rc=$?"
tail -15 $OUT/${TAG}_pytest_wf.log
timeout 600 python -m pytest tests/test_gpu_fingerprint.py tests/test_real_reads.py -q -x > $OUT/${TAG}_pytest_fp.log 2>&1; echo "pytest fingerprint rc=$?"
tail -5 $OUT/${TAG}_pytest_fp.log
timeout 300 python scripts/wavefront_probe.py > $OUT/${TAG}_wavefront_probe.log 2>&1; echo "wf probe rc=$?"
cat $OUT/${TAG}_wavefront_probe.log
timeout 300 python scripts/fp_probe.py > $OUT/${TAG}_fp_probe.log 2>&1; echo "fp probe rc=$?"
tail -8 $OUT/${TAG}_fp_probe.log
