"""Debug: which stage of the fingerprint kernel deviates for one read (clip vs segmentation)."""
import json
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from wdx_testutil import real4000_rows  # noqa: E402

from oracle import wdx_oracle as o  # noqa: E402
from warpdemux_b200.sig_proc import Fingerprinter, FingerprintConfig  # noqa: E402

g = dict(np.load("tests/golden/real4000_rna004_WDX4.npz"))
full = dict(np.load("tests/golden/_local/real4000_adc_rows.npz"))
idx, rows, adc, num = real4000_rows(g, full)
c = json.loads(str(g["cfg"]))
cfg = {k: c[k] for k in ("padding", "outlier_thresh", "min_obs_per_base", "running_stat_width", "num_events", "barcode_num_events")}
i = int(sys.argv[1]) if len(sys.argv) > 1 else 3313
a0, a1 = int(g["bounds"][i, 0]), int(g["bounds"][i, 1])
row = rows[i].copy()
start, stop = max(0, a0 - 100), min(row.size, a1 + 100)
sig = row[start:stop].copy()
med = np.nanmedian(sig)
mad = np.nanmedian(np.abs(sig - med))
lo, hi = med - 5.0 * mad, med + 5.0 * mad
print("slice", start, stop, "med", float(med), "mad", float(mad), "lo", float(lo), "hi", float(hi), type(lo), "has nan", np.isnan(sig).any())
clipped = np.clip(sig, lo, hi)
fp = Fingerprinter(FingerprintConfig(**cfg), device=0)
work = row.copy()[None, :]
b = fp.extract(work, np.array([a0]), np.array([a1]), clip_in_place=True)
gclip = work[0, start:stop]
d = np.flatnonzero(gclip != clipped)
print("clip differs at", d.size, "samples", d[:10], gclip[d[:10]], clipped[d[:10]], sig[d[:10]])
print("gpu min/max", gclip.min(), gclip.max(), "cpu", clipped.min(), clipped.max())
# oracle on the GPU's clipped slice (clip is idempotent if the bounds agree)
st, f, dw, s = o.fingerprint(row, a0, a1, **cfg)
print("oracle == golden", np.array_equal(f, g["fpt"][i]))
print("gpu fpt == golden", np.array_equal(b.fpt[0], g["fpt"][i]), "n diff", (b.fpt[0] != g["fpt"][i]).sum())
# all 111 event means / change points via oracle internals
from scipy.signal import find_peaks
n = clipped.size
w = min(12, round(n / 110)); m_obs = min(6, round(n / 110 / 2))
sc = o.windowed_t_test(clipped, w)
pk, _ = find_peaks(sc, distance=m_obs)
order = np.argsort(sc[pk])
print("n", n, "w", w, "m_obs", m_obs, "peaks", pk.size, "110th/111th best scores", sc[pk][order[-110]], sc[pk][order[-111]] if pk.size > 110 else None)
srt = np.sort(sc[pk])
print("ties among peak scores:", (np.diff(srt) == 0).sum(), "near threshold", srt[-112:-108])
