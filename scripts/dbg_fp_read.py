"""Debug: fingerprint kernel vs golden for single reads of the 4000-read fixture (needs tests/golden/_local)."""
import json
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from wdx_testutil import real4000_rows  # noqa: E402

from warpdemux_b200.sig_proc import Fingerprinter, FingerprintConfig  # noqa: E402

g = dict(np.load("tests/golden/real4000_rna004_WDX4.npz"))
full = dict(np.load("tests/golden/_local/real4000_adc_rows.npz"))
idx, rows, adc, num = real4000_rows(g, full)
c = json.loads(str(g["cfg"]))
cfg = {k: c[k] for k in ("padding", "outlier_thresh", "min_obs_per_base", "running_stat_width", "num_events", "barcode_num_events")}
for cap in (0, 6720, 3968, 4032):
    fp = Fingerprinter(FingerprintConfig(max_slice_len=cap, **cfg), device=0)
    for i in [int(a) for a in sys.argv[1:]] or [3313]:
        sl = slice(i, i + 1)
        b = fp.extract(rows[sl], g["bounds"][sl, 0], g["bounds"][sl, 1])
        print("cap", cap, "read", i, "status", b.status, "fpt equal", np.array_equal(b.fpt[0], g["fpt"][i]), "max diff", np.nanmax(np.abs(b.fpt[0] - g["fpt"][i])))
        print(" dwell equal", np.array_equal(b.dwell[0], g["dwell"][i]), b.dwell[0].tolist(), g["dwell"][i].tolist())
        print(" stats", b.stats[0].tolist(), g["stats"][i].tolist())
    fp.close()
# whole batch of 1000 around it
fp = Fingerprinter(FingerprintConfig(max_slice_len=6720, **cfg), device=0)
sl = slice(3000, 4000)
ok = g["success"][sl].astype(np.uint8)
b = fp.extract(rows[sl], g["bounds"][sl, 0], g["bounds"][sl, 1], detect_ok=ok)
good = g["status"][sl] == 0
bad = np.flatnonzero(good & ~np.all(b.fpt == g["fpt"][sl], axis=1))
print("batch mismatches", (bad + 3000).tolist())
