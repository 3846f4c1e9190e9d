"""Quick device-side throughput probe of the fused DTW+SVC kernel (not the bench)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from warpdemux_b200 import _lib, model_io  # noqa: E402
from warpdemux_b200.device_model import DeviceModel  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "models")


def run(name, mode, n, reps=3):
    m = model_io.load_npz(os.path.join(GOLD, name + ".npz"))
    d = DeviceModel(m, 0)
    d.enable_timing(True)
    rng = np.random.default_rng(0)
    X = m.sv[rng.integers(0, m.n_sv, n)] + 0.35 * rng.standard_normal((n, m.L))
    Xd = torch.from_numpy(X).cuda()
    lab = torch.empty(n, dtype=torch.int64, device="cuda")
    best = 1e30
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.time()
        d.predict_raw(Xd, n, _lib.WDX_F64, _lib.MODES[mode], lab, None, None, None, None, stream=0)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms, nl = d.last_kernel_ms()
        if r > 0:
            best = min(best, ms)
    cells = n * m.n_sv * m.band_cells()
    out = dict(model=name, mode=mode, n=n, kernel_ms=round(best, 3), launches=nl, wall_ms=round(wall * 1e3, 3),
               reads_per_s=round(n / (best * 1e-3)), gcups=round(cells / (best * 1e-3) / 1e9, 1))
    print(json.dumps(out), flush=True)
    d.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for name, mode, n in [("WDX4_rna004_v1_0", "fast", 1 << 20), ("WDX10_rna004_v1_0", "fast", 1 << 19),
                          ("WDX4_rna004_v1_0", "exact", 1 << 18), ("WDX10_rna004_v1_0", "exact", 1 << 17),
                          ("WDX4_rna004_v1_0", "guarded", 1 << 20),
                          ("WDX10_rna004_v1_0", "fast", 1000), ("WDX10_rna004_v1_0", "fast", 512),
                          ("WDX10_rna004_v1_0", "fast", 1)]:
        run(name, mode, n)
