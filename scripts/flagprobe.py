import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from warpdemux_b200 import _lib, model_io
from warpdemux_b200.device_model import DeviceModel
m = model_io.load_npz('/root/repo/tests/golden/models/WDX10_rna004_v1_0.npz')
rng = np.random.default_rng(0)
X = m.sv[rng.integers(0, m.n_sv, 2048)] + 0.35 * rng.standard_normal((2048, m.L))
d = DeviceModel(m, 0); d.enable_timing(True)
for b in (8, 64, 512, 2048):
    lab, prob, conf, flags = d.predict(X[:b], mode="guarded")
    ms_e, nl_e = d.last_kernel_ms_mode(True); ms_f, nl_f = d.last_kernel_ms_mode(False)
    thr = m.thresholds[np.argmax(prob, 1)]
    print(b, "recomputed", int((flags & 2).sum()), "min|conf-thr|", float(np.abs(conf - thr).min()), "min conf", float(conf.min()),
          "exact ms", round(ms_e, 4), nl_e, "fast ms", round(ms_f, 4), nl_f)
