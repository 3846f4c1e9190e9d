"""Device-side throughput of the warp-wavefront DTW kernel (L > 64) through wdx_distance_matrix_to:
device-resident X, Y and output, CUDA events around the call on torch's current stream."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from warpdemux_b200 import _lib  # noqa: E402


def band_cells(L, w):
    if w <= 0 or w > L:
        w = L
    return sum(min(L, i + w) - max(0, i - w + 1) for i in range(L))


def run(L, w, nX, nY, mode, reps=3):
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((nX, L))).cuda()
    Y = torch.from_numpy(rng.standard_normal((nY, L))).cuda()
    out = torch.empty((nX, nY), dtype=torch.float32, device="cuda")
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    best = 1e30
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.wdx_distance_matrix_to(X.data_ptr(), nX, Y.data_ptr(), nY, L, w, 0.1, _lib.MODES[mode], out.data_ptr(),
                                        _lib.WDX_F32, 0, st)
        _lib.check(rc, "wdx_distance_matrix_to")
        e1.record()
        torch.cuda.synchronize()
        if r:
            best = min(best, e0.elapsed_time(e1))
    cells = nX * nY * band_cells(L, w)
    print(json.dumps(dict(L=L, window=w, pairs=nX * nY, mode=mode, ms=round(best, 3),
                          gcups_band=round(cells / (best * 1e-3) / 1e9, 1),
                          gcups_full=round(nX * nY * L * L / (best * 1e-3) / 1e9, 1))), flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for mode in ("fast", "exact"):
        for L, w, nX, nY in [(64, 0, 4096, 2048), (128, 0, 2048, 1024), (128, 16, 2048, 1024), (512, 0, 512, 512),
                             (512, 50, 512, 512), (2048, 0, 128, 128), (2048, 200, 128, 128), (8192, 0, 32, 32)]:
            run(L, w, nX, nY, mode)
    # thread-per-pair generic kernel vs wavefront for short non-specialised shapes
    for thr in ("65", "1"):
        os.environ["WDX_WAVEFRONT_MIN_L"] = thr
        print("WDX_WAVEFRONT_MIN_L=" + thr)
        for mode in ("fast", "exact"):
            for L, w, nX, nY in [(16, 0, 8192, 4096), (32, 0, 8192, 2048), (48, 10, 4096, 2048), (64, 0, 4096, 2048)]:
                run(L, w, nX, nY, mode)
