"""The whole minibatch step (MinibatchDemuxer.run) on DEVICE-RESIDENT rows: CNN -> validation / LLR -> fingerprint -> DTW-SVC
chained through device buffers, results downloaded; with and without the LLR tail overlapped by the fingerprint pass."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from warpdemux_b200 import model_io as _mio  # noqa: E402
from warpdemux_b200.detect import cnn, combined  # noqa: E402
from warpdemux_b200.file_proc import MinibatchDemuxer  # noqa: E402
from warpdemux_b200.models.dtw_svm import DTW_SVM  # noqa: E402

small = _mio.load_npz(os.path.join(ROOT, "tests", "golden", "models", "WDX4_rna004_v1_0.npz"))
ds = bench.chain_dataset()
reps = int(os.environ.get("REPS", "2"))
sig = torch.from_numpy(np.tile(ds["sig"], (reps, 1))).cuda()
lens = np.tile(ds["lens"], reps)
n = sig.shape[0]
md = cnn.load_cnn_model(os.path.join(ROOT, "tests", "golden", "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
out = {"reads": n, "data": ds["kind"]}
labels = {}
for overlap in (False, True):
    mp4 = DTW_SVM(small, device=0, mode="guarded")
    dmx = MinibatchDemuxer(mp4, md, core=cnn.CoreConfig(), cnn_boundaries=cnn.CNNBoundariesConfig(polya_cand_k=5), device=0,
                           llr=combined.LLRConfig(), lanes=1, overlap_llr_tail=overlap)
    for _ in range(2):
        r = dmx.run(sig, lens, return_df=False)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        r = dmx.run(sig, lens, return_df=False)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    labels[overlap] = r.labels.copy()
    out["overlap_llr_tail" if overlap else "sequential"] = {"ms": round(best * 1e3, 3), "reads_per_s": round(n / best)}
    dmx.close()
out["labels_identical"] = bool(np.array_equal(labels[False], labels[True]))
print(json.dumps(out))
