#!/usr/bin/env bash
# multi-GPU visit: usage scripts/gpu_multi.sh TAG N  -- the 2-GPU test (N >= 2) and bench under torchrun at N ranks
set -u
TAG=${1:-r2mg}; N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_real_reads.py tests/test_sharding_gloo.py -q -x -k "second_device or shard" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 \
   --no-cpu-baseline --no-extra-modes --config2-reads 1000000 --cadence-seconds 0 --single-call-reads 0 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err ) 2> $OUT/${TAG}_bench_n$N.time
tail -3 $OUT/${TAG}_bench_n$N.err; cat $OUT/${TAG}_bench_n$N.time
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["e2e"]["value"])
print("sharded:", json.dumps(d.get("e2e_sharded_with_label_gather"))[:400])
print("chain stream:", json.dumps(d.get("raw_signal_chain_stream"))[:600])
PY
