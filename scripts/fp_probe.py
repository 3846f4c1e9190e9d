"""Device-side throughput probe of the fingerprint kernel (not the bench)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from wdx_testutil import synth_adapter_signals  # noqa: E402
from warpdemux_b200.sig_proc import Fingerprinter  # noqa: E402


def main(n_base=512, reps=int(os.environ.get("FP_REPS", "64"))):
    okd = None
    if os.environ.get("FP_REAL"):   # the 4000 real reads (rows of bench.chain_dataset) with the reference's own boundaries
        import bench
        ds = bench.chain_dataset()
        with np.load(os.path.join(ROOT, "tests", "golden", "real4000_rna004_WDX4.npz")) as z:
            bounds, success = z["bounds"], z["success"]
        sig = ds["sig"]
        a0, a1 = bounds[:, 0].astype(np.int64).copy(), bounds[:, 1].astype(np.int64).copy()
        a0[success == 0] = 0
        a1[success == 0] = 0
        n_base, reps = sig.shape[0], max(1, reps // 8)
        okd = torch.from_numpy(np.tile(success.astype(np.uint8), reps)).cuda()
        os.environ.setdefault("FP_MAX_SLICE", "6720")
    else:
        sig, a0, a1 = synth_adapter_signals(n_base, seed=21, width=9000)
    short = int(os.environ.get("FP_SHORT", "0"))
    if short:   # adapters cut to `short` samples (CTA-shape experiments: more CTAs fit an SM with smaller slices)
        a1 = np.minimum(a1, a0 + short)
    n = n_base * reps
    sd = torch.from_numpy(sig).cuda().repeat(reps, 1).contiguous()
    a0d = torch.from_numpy(a0).cuda().repeat(reps).contiguous()
    a1d = torch.from_numpy(a1).cuda().repeat(reps).contiguous()
    lens = (~np.isnan(sig)).sum(axis=1)
    sl = np.minimum(lens, a1 + 100) - np.maximum(0, a0 - 100)
    bytes_alg = int(sl.sum()) * 4 * reps + n * 25 * 8
    fpt = torch.empty((n, 25), dtype=torch.float64, device="cuda")
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    from warpdemux_b200.sig_proc import FingerprintConfig
    fp = Fingerprinter(FingerprintConfig(max_slice_len=int(os.environ.get("FP_MAX_SLICE", "7040"))), device=0)
    fp.enable_timing(True)
    stream = torch.cuda.current_stream().cuda_stream
    best = 1e30
    for r in range(5):
        fp.extract_raw(sd, n, sig.shape[1], a0d, a1d, fpt, st, detect_ok=okd, stream=stream)
        torch.cuda.synchronize()
        ms, nl = fp.last_kernel_ms()
        if r:
            best = min(best, ms)
    ok = int((st == 0).sum().item())
    from warpdemux_b200 import _lib
    L = _lib.load()
    if hasattr(L, "wdx_fp_prof_dump"):   # -DWDX_FP_PROF build: cycles per phase and read (thread 0 of every CTA)
        import ctypes as C
        buf = (C.c_uint64 * 32)()
        L.wdx_fp_prof_dump(None, 1)
        fp.extract_raw(sd, n, sig.shape[1], a0d, a1d, fpt, st, detect_ok=okd, stream=stream)
        torch.cuda.synchronize()
        L.wdx_fp_prof_dump(buf, 0)
        names = ["load", "medians", "clip", "ttest", "localmax", "nbr_sets", "rounds", "-", "scan+list", "topk", "means", "normalize", "out", "n_mean", "n_ss", "n_sqrt"]
        ph = {nm: round(buf[i] / n) for i, nm in enumerate(names) if nm != "-"}
        ph["sum"] = sum(ph.values())
        print(json.dumps({"cycles_per_read": ph}), flush=True)
    print(json.dumps(dict(reads=n, ok=ok, mean_slice=float(sl.mean()), kernel_ms=round(best, 3), launches=nl,
                          reads_per_s=round(n / best * 1e3), alg_GBps=round(bytes_alg / best / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
