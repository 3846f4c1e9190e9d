#!/usr/bin/env bash
# usage: scripts/gpu_tests.sh TAG [pytest args...]   (default: the whole -m gpu suite)
set -u
TAG=${1:-r2t}; shift || true
OUT=gpurun_out
mkdir -p $OUT
if [ $# -eq 0 ]; then set -- tests; fi
timeout 1700 python -m pytest "$@" -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
