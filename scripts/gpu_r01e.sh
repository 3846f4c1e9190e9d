#!/usr/bin/env bash
# GPU visit r01e: everything in gpu_quick.sh + boundary-CNN probe + ncu captures of the CNN kernels.
set -u
TAG=${1:-r01e}
OUT=gpurun_out
bash scripts/gpu_quick.sh $TAG
timeout 300 python scripts/cnn_probe.py > $OUT/${TAG}_cnn_probe.log 2>&1; echo "cnn probe rc=$?"
cat $OUT/${TAG}_cnn_probe.log | tail -5
CNN_PROBE_READS=1024 CNN_PROBE_MODES=fast timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/${TAG}_cnn_launches.csv python scripts/cnn_probe.py > $OUT/${TAG}_cnn_under_ncu.log 2>&1
tail -3 $OUT/${TAG}_cnn_launches.csv
CNN_PROBE_READS=1024 CNN_PROBE_MODES=fast timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnn_tc -s 1 -c 1 \
    -o $OUT/${TAG}_cnnprof -f python scripts/cnn_probe.py > $OUT/${TAG}_cnnprof.log 2>&1
tail -2 $OUT/${TAG}_cnnprof.log
ls -la $OUT | tail
