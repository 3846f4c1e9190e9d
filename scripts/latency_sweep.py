"""Small-batch latency vs the number of SV ranges (sv_splits) of the fused kernel (not the bench)."""
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

name = os.environ.get("MODEL", "WDX10_rna004_v1_0")
m = model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", name + ".npz"))
rng = np.random.default_rng(0)
X = m.sv[rng.integers(0, m.n_sv, 2048)] + 0.35 * rng.standard_normal((2048, m.L))
mdl = DTW_SVM(m, device=0, mode="guarded")
dm = mdl._device_model()
Xd = torch.from_numpy(X).cuda()
lab = torch.empty(2048, dtype=torch.int64, device="cuda")
conf = torch.empty(2048, dtype=torch.float64, device="cuda")
prob = torch.empty((2048, m.k), dtype=torch.float64, device="cuda")
for mode in ("fast", "exact", "guarded"):
    MODE = _lib.MODES[mode]
    for b in (1, 8, 64, 512):
        row = {}
        for splits in (0, 64, 128, 256, 512, 1024, m.n_sv):
            _lib.check(_lib.load().wdx_model_set_sv_splits(dm._h, splits), "splits")
            fn = lambda: (dm.predict_raw(Xd, b, _lib.WDX_F64, MODE, lab, conf, prob, None, None, stream=0), torch.cuda.synchronize())
            for _ in range(20):
                fn()
            ts = []
            for _ in range(100):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            row[splits] = round(float(np.percentile(np.array(ts) * 1e3, 50)), 4)
        print(json.dumps(dict(model=name, mode=mode, batch=b, p50_ms_by_splits=row)), flush=True)
