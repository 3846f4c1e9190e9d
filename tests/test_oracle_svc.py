"""Pins the libsvm restatement (oracle/wdx_oracle.c) against the real
sklearn/libsvm binary: live on a freshly fitted model, and through the golden
outputs of the reference's DTW_SVM.predict on the shipped models."""
import warnings

import numpy as np
import pytest

from oracle import wdx_oracle as o
from warpdemux_b200 import model_io


def _fit_toy(k=4, n_per=40, L=25, seed=0):
    from sklearn.svm import SVC

    rng = np.random.default_rng(seed)
    templ = rng.standard_normal((k, L))
    y = np.repeat(np.arange(k), n_per)
    X = templ[y] + 0.8 * rng.standard_normal((k * n_per, L))
    D = o.distance_matrix_to(X, X, 15, 0.1)
    K = o.pdist_kernel(D, 1.0, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        svc = SVC(kernel="precomputed", probability=True, C=1.0, class_weight="balanced", random_state=0)
        svc.fit(K.astype(np.float64), y)
    return svc, X, y


def test_libsvm_restatement_matches_sklearn_live():
    svc, X, y = _fit_toy()
    k = 4

    class Ref:  # shaped like the reference DTW_SVM (dtw_base.py:13-25)
        model = svc
        _X = X
        label_mapper = {0: 3, 1: 4, 2: 5, 3: -1}
        thresholds = np.array([0.2, 0.3, 0.1, 0.0])
        window, penalty, gamma, pwr_dist, block_size, noise_class = 15, 0.1, 1.0, 1, 500, True

    m = model_io.from_reference_model(Ref, name="toy")
    assert m.n_sv == svc.support_.size and m.k == k
    rng = np.random.default_rng(1)
    Xq = X[rng.integers(0, X.shape[0], 64)] + 0.5 * rng.standard_normal((64, X.shape[1]))
    Kfull = o.pdist_kernel(o.distance_matrix_to(Xq, X, 15, 0.1), 1.0, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = svc.predict_proba(Kfull)
        want_dec = svc.decision_function(Kfull) if False else None
    Ksv = o.pdist_kernel(o.distance_matrix_to(Xq, m.sv, 15, 0.1), 1.0, 1)
    got, _ = o.svc_predict_proba(Ksv, m)
    # same arithmetic, same order: differences only from libm exp in the sigmoid
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)
    assert np.array_equal(got.argmax(1), want.argmax(1))


def test_golden_predict(models, golden_predict):
    """oracle.predict == reference DTW_SVM.predict (real sklearn) on the shipped models."""
    for name, g in golden_predict.items():
        m = models[name]
        pred, prob, conf, _ = o.predict(m, g["X"])
        assert np.array_equal(pred, g["y_pred"]), name
        np.testing.assert_allclose(prob, g["y_prob"], rtol=0, atol=1e-13)
        # the DataFrame view (models/utils.py:36-43)
        cols = list(g["df_columns"])
        assert cols[0] == "predicted_barcode" and cols[1] == "confidence_score"
        assert np.array_equal(g["df_values"][:, 0].astype(np.int64), pred)
        np.testing.assert_allclose(g["df_values"][:, 1], conf.round(3), atol=1e-12)


def test_predict_c_path_agrees(models, golden_predict):
    """The single-C-call timed arm differs from the numpy-level oracle only in
    float32 exp (libm expf vs numpy SIMD): probabilities to ~1e-6, labels equal
    on this set."""
    name = "WDX4_rna004_v1_0"
    m, g = models[name], golden_predict[name]
    pred, prob, conf = o.predict_c(m, g["X"])
    np.testing.assert_allclose(prob, g["y_prob"], rtol=0, atol=2e-6)
    assert np.array_equal(pred, g["y_pred"])
    pred2, prob2, _ = o.predict_threaded(m, g["X"], threads=2, minibatch=100)
    assert np.array_equal(pred2, pred) and np.array_equal(prob2, prob)


def test_error_behaviour(models):
    m = models["WDX4_rna004_v1_0"]
    with pytest.raises(ValueError, match="same number of columns"):
        o.predict(m, np.zeros((3, 24)))
