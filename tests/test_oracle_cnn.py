"""Oracle of the boundary CNN (oracle/wdx_oracle_cnn.py) against the golden vectors produced by the
reference's own adapted.detect.cnn (oracle/make_golden_cnn.py, real reads)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from wdx_testutil import cnn_golden_signals


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "cnn_detect_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def weights():
    from oracle import wdx_oracle_cnn as oc

    return oc.load_weights_npz(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"))


def _cfg(gold):
    from oracle import wdx_oracle_cnn as oc

    return oc.CnnConfig(**json.loads(str(gold["cfg"])))


def test_prepare_data_bit_identical(gold):
    from oracle import wdx_oracle_cnn as oc

    x = oc.prepare_data(cnn_golden_signals(gold), _cfg(gold))
    assert x.dtype == np.float32
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(gold["x_sha256"])
    assert np.array_equal(x[: gold["x_rows"].shape[0]], gold["x_rows"])


def test_scores_match_reference_forward(gold, weights):
    from oracle import wdx_oracle_cnn as oc

    s = oc.cnn_scores(gold["x_rows"], weights)
    assert s.shape == gold["scores"].shape
    # same torch build, same convolution primitives: float32 summation-order noise at most
    assert np.abs(s - gold["scores"]).max() <= 1e-4 * max(1.0, np.abs(gold["scores"]).max())


def test_boundaries_identical(gold, weights):
    from oracle import wdx_oracle_cnn as oc

    preds = oc.cnn_detect(cnn_golden_signals(gold), weights, _cfg(gold))
    assert np.array_equal(preds, gold["preds"])


def test_predict_group_shift_quirk():
    """A read without a single peak shifts the candidate rows of all later reads (cnn.py:147-158)."""
    from oracle import wdx_oracle_cnn as oc

    cfg = oc.CnnConfig(min_obs_adapter=0, max_obs_adapter=400, downscale_factor=10, polya_cand_k=3)
    T = 60
    s = np.full((3, 2, T), -9.0, dtype=np.float32)
    s[:, 0, 5] = 1.0                     # adapter end at 5 for every read
    s[0, 1, 20], s[0, 1, 30] = 2.0, 3.0   # read 0: two peaks
    # read 1: everything below SCORE_EXCL -> argmax inside the unmasked range, but no strict maximum
    s[2, 1, 40] = 4.0                     # read 2: one peak
    out = oc.cnn_predict(s, cfg)
    assert out[0].tolist() == [5, 30, 20, 0]
    assert out[:, 0].tolist() == [5, 5, 5]
    # what row 1 / row 2 hold is decided by the flattened find_peaks; just pin it against itself
    assert out.shape == (3, 4)
