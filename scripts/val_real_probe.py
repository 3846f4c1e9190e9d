"""Validation (+ LLR) kernel timing on the REAL reads of the chain data set (bench.chain_dataset), device buffers.
WDX_B200_LIB selects a kernel variant.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from warpdemux_b200.detect import cnn, combined  # noqa: E402

ds = bench.chain_dataset()
sig, lens = ds["sig"], ds["lens"]
reps = max(1, 8000 // sig.shape[0])
n, stride, k = sig.shape[0] * reps, sig.shape[1], 5
d_sig = torch.from_numpy(np.tile(sig, (reps, 1))).cuda()
d_len = torch.from_numpy(np.tile(lens, reps)).cuda()
model = cnn.load_cnn_model(os.path.join(ROOT, "tests", "golden", "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
d_preds = torch.zeros((n, 1 + k), dtype=torch.int64, device="cuda")
side = torch.cuda.Stream()
sp = side.cuda_stream
out = {"lib": os.environ.get("WDX_B200_LIB", "default"), "n": n, "data": ds["kind"][:20]}
with torch.cuda.stream(side):
    cnn.detect_raw(model, cnn.CoreConfig(), k, d_sig, n, stride, d_preds, stream=sp)
    side.synchronize()
    for name, llr, vo in (("verdict_only", None, True), ("full_report", None, False), ("verdict_only_llr", combined.LLRConfig(), True)):
        v = combined.Validator(combined.ValidateConfig(), device=0, verdict_only=vo, llr=llr)
        d_suc = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d_info = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        d_bounds = torch.zeros((n, 3), dtype=torch.int64, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = None
        for it in range(5):
            ev[0].record()
            v.run_raw(d_sig, n, stride, d_len, d_preds, 1 + k, d_suc, d_info, d_bounds, None, stream=sp)
            ev[1].record()
            side.synchronize()
            t = ev[0].elapsed_time(ev[1])
            best = t if best is None or (it and t < best) else best
        out[name + "_ms"] = round(best, 4)
        from warpdemux_b200 import _lib
        L = _lib.load()
        if name == "verdict_only" and hasattr(L, "wdx_validate_prof_dump"):   # -DWDX_FP_PROF build: cycles per phase and read
            import ctypes as C
            buf = (C.c_uint64 * 32)()
            L.wdx_validate_prof_dump(None, 1)
            v.run_raw(d_sig, n, stride, d_len, d_preds, 1 + k, d_suc, d_info, d_bounds, None, stream=sp)
            side.synchronize()
            L.wdx_validate_prof_dump(buf, 0)
            names = ["load", "1a", "1b", "focus", "1c", "A_hist", "A_scan", "A_gather", "A_rank", "A_results", "B_hist", "B_scan", "B_gather",
                     "B_rank+setup", "verdict", "rest"]
            ph = {nm: round(buf[i] / n) for i, nm in enumerate(names)}
            ph["sum"] = sum(ph.values())
            out["cycles_per_read"] = ph
        if name == "verdict_only_llr" and hasattr(L, "wdx_validate_prof_dump"):   # slots 16.. : the LLR kernel, per read that reaches it
            import ctypes as C
            buf = (C.c_uint64 * 32)()
            L.wdx_validate_prof_dump(None, 1)
            v.run_raw(d_sig, n, stride, d_len, d_preds, 1 + k, d_suc, d_info, d_bounds, None, stream=sp)
            side.synchronize()
            L.wdx_validate_prof_dump(buf, 0)
            lnames = ["row+medmad", "downscale", "cumsum", "gains+start_end", "nanstd", "find_peaks", "plateau", "split_peak", "gains2+polya", "rest"]
            nl = max(1, int(((d_info[:, 3] & 12) != 0).sum().item()))
            out["llr_cycles_per_llr_read_both_stages"] = {nm: round(buf[16 + i] / nl) for i, nm in enumerate(lnames)}
        out[name + "_ok"] = int(d_suc.sum().item())
        out[name + "_chk"] = int((d_bounds.sum() + d_info[:, 0].sum()).item())
        v.close()
print(json.dumps(out))
