"""A few batch-B predict calls on device buffers (for an ncu launch list)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from warpdemux_b200 import _lib, model_io  # noqa: E402
from warpdemux_b200.models.dtw_svm import DTW_SVM  # noqa: E402

m = model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", "WDX10_rna004_v1_0.npz"))
rng = np.random.default_rng(0)
X = m.sv[rng.integers(0, m.n_sv, 2048)] + 0.35 * rng.standard_normal((2048, m.L))
mode = os.environ.get("MODE", "fast")
b = int(os.environ.get("BATCH", "1"))
mdl = DTW_SVM(m, device=0, mode=mode)
dm = mdl._device_model()
Xd = torch.from_numpy(X).cuda()
lab = torch.empty(2048, dtype=torch.int64, device="cuda")
conf = torch.empty(2048, dtype=torch.float64, device="cuda")
prob = torch.empty((2048, m.k), dtype=torch.float64, device="cuda")
for _ in range(4):
    dm.predict_raw(Xd, b, _lib.WDX_F64, _lib.MODES[mode], lab, conf, prob, None, None, stream=0)
    torch.cuda.synchronize()
