"""Per-call latency of the predict path for live-sized batches (not the bench)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from warpdemux_b200 import _lib, model_io  # noqa: E402
from warpdemux_b200.models.dtw_svm import DTW_SVM  # noqa: E402

m = model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", "WDX10_rna004_v1_0.npz"))
rng = np.random.default_rng(0)
X = m.sv[rng.integers(0, m.n_sv, 2048)] + 0.35 * rng.standard_normal((2048, m.L))
mdl = DTW_SVM(m, device=0, mode=os.environ.get("MODE", "guarded"))
dm = mdl._device_model()
Xd = torch.from_numpy(X).cuda()
lab = torch.empty(2048, dtype=torch.int64, device="cuda")
conf = torch.empty(2048, dtype=torch.float64, device="cuda")
prob = torch.empty((2048, m.k), dtype=torch.float64, device="cuda")
MODE = _lib.MODES[mdl.mode]
for b in (1, 8, 64, 512, 2048):
    for name, fn in (("numpy_api", lambda: mdl.predict(X[:b], nproc=1)),
                     ("raw_device", lambda: (dm.predict_raw(Xd, b, _lib.WDX_F64, MODE, lab, conf, prob, None, None, stream=0),
                                             torch.cuda.synchronize()))):
        for _ in range(30):
            fn()
        ts = []
        for _ in range(300):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(json.dumps(dict(batch=b, path=name, mode=mdl.mode, p50_ms=round(float(np.percentile(ts, 50)), 4),
                              p99_ms=round(float(np.percentile(ts, 99)), 4))), flush=True)
