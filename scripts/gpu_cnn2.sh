#!/usr/bin/env bash
# boundary-CNN iteration: parity tests of the CNN + the chain on real reads, then the chain timing probe
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_cnn.py tests/test_real_reads.py tests/test_gpu_llr.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log | cut -c1-1500
timeout 300 python scripts/chain_probe.py 2>&1 | tail -1 > $OUT/${TAG}_chain.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_chain.json"))
print(json.dumps(d["device_resident"])); print(json.dumps(d["llr_fallback"]))
print("run", d["e2e"]["reads_per_s"], "stream", d["e2e_pipelined_minibatches"]["reads_per_s"], "stream adc", d["e2e_pipelined_minibatches_adc"]["reads_per_s"])
PY
