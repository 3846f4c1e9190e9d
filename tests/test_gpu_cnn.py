"""GPU parity of the boundary-CNN stage (wdx_cnn_* through the C ABI) against the oracle
(oracle/wdx_oracle_cnn.py) and the golden vectors of the reference's own adapted.detect.cnn."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from wdx_testutil import cnn_golden_signals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "cnn_detect_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ocfg(gold):
    from oracle import wdx_oracle_cnn as oc

    return oc.CnnConfig(**json.loads(str(gold["cfg"])))


@pytest.fixture(scope="module")
def weights():
    from oracle import wdx_oracle_cnn as oc

    return oc.load_weights_npz(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"))


@pytest.fixture(scope="module")
def model():
    from warpdemux_b200.detect import cnn

    m = cnn.load_cnn_model(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
    yield m
    m.close()


def _cfgs(ocfg):
    from warpdemux_b200.detect import cnn

    core = cnn.CoreConfig(min_obs_adapter=ocfg.min_obs_adapter, max_obs_adapter=ocfg.max_obs_adapter,
                          downscale_factor=ocfg.downscale_factor)
    return cnn.CNNBoundariesConfig(polya_cand_k=ocfg.polya_cand_k), core


def _edge_rows(m, rng):
    rows = np.full((6, m), np.nan, dtype=np.float32)
    rows[0, :] = np.nan                                             # empty read
    rows[1, :1500] = 80 + 10 * rng.standard_normal(1500)            # ends inside the first downscale blocks
    rows[2, :] = 75.0                                               # constant: MAD = 0 -> 0/0
    rows[3, :] = 80 + 10 * rng.standard_normal(m)
    rows[3, 5000:5003] = np.nan                                     # NaN hole
    rows[4, :1005] = 90.0                                           # 5 valid samples after min_obs_adapter
    rows[5, :] = 80 + 10 * rng.standard_normal(m)
    rows[5, 4000] = np.inf
    return rows


def test_prepare_bit_identical(gold, ocfg, model):
    from oracle import wdx_oracle_cnn as oc
    from warpdemux_b200.detect import cnn

    _, core = _cfgs(ocfg)
    sig = cnn_golden_signals(gold)
    sig = np.concatenate([sig, _edge_rows(sig.shape[1], np.random.default_rng(3))])
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = oc.prepare_data(sig, ocfg)
    got = cnn.prepare_data(sig, core, model, k=ocfg.polya_cand_k)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a row stride that is not a multiple of the factor: zero padding of the last block (downscale.py:22-27)
    odd = np.ascontiguousarray(sig[:8, :18493])
    with np.errstate(all="ignore"):
        want = oc.prepare_data(odd, ocfg)
    got = cnn.prepare_data(odd, core, model, k=ocfg.polya_cand_k)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _crafted_scores(rng, n, T):
    # continuous values: two DISTINCT peaks of exactly equal height closer than `distance` would be resolved by
    # numpy's unstable argsort inside scipy (unspecified); plateaus (ties inside ONE peak) are well defined
    s = (rng.standard_normal((n, 2, T)) * 6).astype(np.float32)
    for r in range(16, n):                       # short plateaus sprinkled over the later reads
        for p0 in rng.integers(5, T - 10, size=30):
            s[r, 1, p0:p0 + int(rng.integers(2, 5))] = s[r, 1, p0]
    s[3, 1, :] = -9.0                            # nothing above SCORE_EXCL: read without a peak (row-shift quirk)
    s[7, 1, :] = -5.0                            # a read that is one plateau
    s[10, 1, T - 1] = 50.0                       # maximum on the last sample, next read starts high
    s[11, 0, 0] = 60.0                           # adapter end 0 -> no leading mask
    s[11, 1, 0:3] = 40.0
    s[12, 1, 100:140] = 7.0                      # long plateau
    s[13, 1, :] = np.linspace(-3, 3, T)          # monotone
    s[14, 0, :] = -1.0                           # flat channel 0: argmax 0
    return s.astype(np.float32)


@pytest.mark.parametrize("k", [5, 10])
def test_predict_matches_oracle_on_crafted_scores(ocfg, model, k):
    from oracle import wdx_oracle_cnn as oc
    from warpdemux_b200.detect import cnn

    rng = np.random.default_rng(11)
    params, core = _cfgs(ocfg)
    params.polya_cand_k = k
    for T in (1750, 97):
        s = _crafted_scores(rng, 40, T) if T > 200 else (rng.standard_normal((33, 2, T)) * 3).astype(np.float32)
        cfg = oc.CnnConfig(min_obs_adapter=ocfg.min_obs_adapter, max_obs_adapter=ocfg.max_obs_adapter,
                           downscale_factor=ocfg.downscale_factor, polya_cand_k=k)
        want = oc.cnn_predict(s, cfg)
        got, flags = cnn.cnn_predict(s, model, params, core, return_flags=True)
        assert not (flags & cnn.FLAG_CHAIN).any()
        assert np.array_equal(got, want), np.argwhere(got != want)[:5]


def test_exact_mode_matches_golden(gold, ocfg, model):
    from warpdemux_b200.detect import cnn

    params, core = _cfgs(ocfg)
    sig = cnn_golden_signals(gold)
    scores, preds = cnn.cnn_score_batch(sig, model, params, core, mode="exact")
    nrow = gold["scores"].shape[0]
    assert scores.shape[1:] == gold["scores"].shape[1:]
    tol = 2e-5 * max(1.0, float(np.abs(gold["scores"]).max()))   # float32 summation order vs torch's CPU convolution
    assert np.abs(scores[:nrow] - gold["scores"]).max() <= tol
    assert np.array_equal(preds, gold["preds"])
    assert np.array_equal(cnn.cnn_detect(sig, model, params, core, mode="exact"), gold["preds"])


def test_fast_mode_matches_exact_and_golden(gold, ocfg, model):
    from warpdemux_b200.detect import cnn

    params, core = _cfgs(ocfg)
    sig = cnn_golden_signals(gold)
    se, pe = cnn.cnn_score_batch(sig, model, params, core, mode="exact")
    sf, pf = cnn.cnn_score_batch(sig, model, params, core, mode="fast")
    scale = max(1.0, float(np.abs(se).max()))
    err = float(np.abs(sf - se).max())
    print("fast vs exact max abs score error", err, "relative to max |score|", err / scale)
    assert err <= 2e-5 * scale      # fp16 hi/lo split, three products: float32-class accuracy
    assert np.array_equal(pf, gold["preds"])
    pg, flags = cnn.cnn_detect(sig, model, params, core, mode="guarded", return_flags=True)
    assert np.array_equal(pg, gold["preds"])
    assert not (flags & (cnn.FLAG_CHAIN | cnn.FLAG_RANGE)).any()


def test_batch_of_perturbed_reads_all_modes_agree_with_oracle(gold, ocfg, model, weights):
    """A larger flattened batch (several chunks): real reads + noise; oracle = numpy prepare + torch fp32 + scipy."""
    from oracle import wdx_oracle_cnn as oc
    from warpdemux_b200.detect import cnn

    params, core = _cfgs(ocfg)
    base = cnn_golden_signals(gold)
    rng = np.random.default_rng(5)
    reps = 40                                                    # 2560 reads > 2 chunks of 1024
    sig = np.concatenate([base + rng.standard_normal(base.shape).astype(np.float32) * np.float32(0.3 + 0.05 * r)
                          for r in range(reps)])
    want = oc.cnn_detect(sig[:384], weights, ocfg)               # CPU: keep it to a few seconds
    pe = cnn.cnn_detect(sig, model, params, core, mode="exact")
    pf = cnn.cnn_detect(sig, model, params, core, mode="fast")
    pg, flags = cnn.cnn_detect(sig, model, params, core, mode="guarded", return_flags=True)
    # the flattened peak search only couples neighbouring reads through the last/first samples, which are masked
    # here, so the first 384 rows of the big batch equal the 384-read oracle batch
    mism_exact = int((pe[:384] != want).any(axis=1).sum())
    mism_fast = int((pf != pe).any(axis=1).sum())
    mism_guard = int((pg != pe).any(axis=1).sum())
    print("rows differing: exact vs oracle", mism_exact, "/384; fast vs exact", mism_fast, "; guarded vs exact", mism_guard,
          "of", len(sig), "; recomputed", int((flags & cnn.FLAG_RECOMPUTED).astype(bool).sum()))
    assert mism_exact == 0
    assert mism_guard == 0
    assert mism_fast <= 2                                        # near-ties only; GUARDED removes them


def test_error_conventions(model, ocfg):
    from warpdemux_b200.detect import cnn

    params, core = _cfgs(ocfg)
    with pytest.raises(ValueError):
        cnn.cnn_detect(np.zeros((2, 900), dtype=np.float32), model, params, core)      # stride <= min_obs_adapter
    with pytest.raises(ValueError):
        cnn.cnn_detect(np.zeros(18500, dtype=np.float32), model, params, core)          # 1-D
    assert cnn.cnn_detect(np.zeros((0, 18500), dtype=np.float32), model, params, core).shape == (0, 1 + params.polya_cand_k)
