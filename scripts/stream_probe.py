"""Production-shape stream (1000-read AdcBatch minibatches from pinned memory) through MinibatchDemuxer.stream:
wall time per minibatch.  Under `ncu --metrics gpu__time_duration.sum` the launch list gives the device time per minibatch."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

n_mb = int(os.environ.get("N_MB", "32"))
lanes = int(os.environ.get("LANES", "2"))
dmx, md, mbs = bench.raw_chain_stream_rank(0, minibatches=n_mb, lanes=lanes)
torch.cuda.synchronize()
out = {"lanes": lanes}
for rep in range(3):
    t0 = time.perf_counter()
    got = sum(int(r.labels.size) for r in dmx.stream(mbs, return_df=False))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"stream_ms_per_minibatch_{rep}"] = round(dt / n_mb * 1e3, 4)
# one minibatch at a time (no overlap): the latency of the chain
t0 = time.perf_counter()
for mb in mbs[:8]:
    dmx.run(mb[0], mb[1], return_df=False)
out["run_ms_per_minibatch"] = round((time.perf_counter() - t0) / 8 * 1e3, 4)
out["reads_per_s"] = got / dt
print(json.dumps(out))
