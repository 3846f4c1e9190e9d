"""Two 1000-read minibatches through MinibatchDemuxer.run (for an ncu launch list)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch  # noqa: E402

from validate_probe import synth_reads  # noqa: E402
from warpdemux_b200 import model_io  # noqa: E402
from warpdemux_b200.detect import cnn  # noqa: E402
from warpdemux_b200.file_proc import MinibatchDemuxer  # noqa: E402
from warpdemux_b200.models.dtw_svm import DTW_SVM  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
sig, lens, _, _ = synth_reads(250, 11500)
sig, lens = np.tile(sig, (4, 1)), np.tile(lens, 4)
model = cnn.load_cnn_model(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
mdl = DTW_SVM(model_io.load_npz(os.path.join(GOLD, "models", "WDX4_rna004_v1_0.npz")), device=0, mode="guarded")
dmx = MinibatchDemuxer(mdl, model, core=cnn.CoreConfig(), cnn_boundaries=cnn.CNNBoundariesConfig(polya_cand_k=5), device=0)
h = torch.from_numpy(sig).pin_memory()
for _ in range(3):
    r = dmx.run(h, lens, return_df=False)
print(int(r.detect_success.sum()), int((r.fp_status == 0).sum()))
