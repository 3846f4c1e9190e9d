"""Throughput probe of the boundary-CNN stage on device-resident signals (reads/s per mode, kernel ms)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from wdx_testutil import cnn_golden_signals  # noqa: E402
from warpdemux_b200.detect import cnn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
with np.load(os.path.join(GOLD, "cnn_detect_rna004.npz")) as z:
    gold = {k: z[k] for k in z.files}
base = cnn_golden_signals(gold)
n = int(os.environ.get("CNN_PROBE_READS", "4096"))
rng = np.random.default_rng(0)
sig = np.concatenate([base + rng.standard_normal(base.shape).astype(np.float32) * np.float32(0.5)
                      for _ in range(n // base.shape[0])])
n = sig.shape[0]
sd = torch.from_numpy(sig).cuda()
model = cnn.load_cnn_model(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
params, core = cnn.CNNBoundariesConfig(polya_cand_k=5), cnn.CoreConfig()
k = 5
preds = torch.zeros((n, 1 + k), dtype=torch.int64, device="cuda")
flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
cnn.enable_timing(model, core, k, True)
stream = torch.cuda.current_stream().cuda_stream
out = {}
for mode in os.environ.get("CNN_PROBE_MODES", "exact,fast,guarded").split(","):
    for _ in range(2):
        cnn.detect_raw(model, core, k, sd, n, sig.shape[1], preds, flags=flags, mode=mode, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        cnn.detect_raw(model, core, k, sd, n, sig.shape[1], preds, flags=flags, mode=mode, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kms, kl = cnn.last_kernel_ms(model, core, k)
    flops = n * 2 * 584 * 64 * 64 * 7 * 2
    out[mode] = {"reads": n, "ms": ms, "reads_per_s": n / (ms * 1e-3), "conv_kernel_ms": kms, "conv_launches": kl,
                 "conv_tflops_fp32_equiv": flops / (kms * 1e-3) / 1e12,
                 "recomputed": int((flags & 2).bool().sum().item())}
    print(json.dumps({mode: out[mode]}), flush=True)
