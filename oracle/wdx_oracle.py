"""CPU oracle for the WarpDemuX classification hot path — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
The product (warpdemux_b200/) never does.

Layers (each cites the reference lines it restates):
  * ctypes bindings of oracle/wdx_oracle.c  (DTW, libsvm probability, decision)
  * numpy-level restatement of DTW_SVM.predict       (warpdemux/models/dtw_svm.py:54-98)
  * numpy/scipy-level restatement of the fingerprint (warpdemux/sig_proc.py:394-605)

Parity status: see the header of wdx_oracle.c.  DTW is a restatement of the
absent dtaidistance 2.3.13, pinned by the KKT known-answer test and by golden
fixtures produced with the reference's own Python + the real sklearn binary.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwdx_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "wdx_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.wdx_oracle_dtw_distance.restype = C.c_double
        L.wdx_oracle_dtw_distance.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double]
        L.wdx_oracle_dtw_matrix.restype = None
        L.wdx_oracle_dtw_matrix.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_int, C.c_int, C.c_double, _dp]
        L.wdx_oracle_dtw_matrix_f32.restype = None
        L.wdx_oracle_dtw_matrix_f32.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_int, C.c_int, C.c_double, _fp]
        L.wdx_oracle_svc_predict_proba.restype = C.c_int
        L.wdx_oracle_svc_predict_proba.argtypes = [_fp, C.c_int64, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp]
        L.wdx_oracle_process_probs.restype = None
        L.wdx_oracle_process_probs.argtypes = [_dp, C.c_int64, C.c_int, _lp, _dp, _lp, _dp]
        L.wdx_oracle_predict.restype = C.c_int
        L.wdx_oracle_predict.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, _lp, _dp, _lp, _dp, _dp]
        L.wdx_oracle_windowed_t_test.restype = None
        L.wdx_oracle_windowed_t_test.argtypes = [_dp, C.c_int64, C.c_int64, _dp]
        L.wdx_oracle_new_means.restype = None
        L.wdx_oracle_new_means.argtypes = [_dp, _lp, C.c_int64, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------
# DTW
# --------------------------------------------------------------------------
def dtw_distance(s1, s2, window: int, penalty: float) -> float:
    s1, s2 = _c64(s1), _c64(s2)
    return float(lib().wdx_oracle_dtw_distance(_d(s1), s1.size, _d(s2), s2.size, int(window or 0), float(penalty or 0.0)))


def dtw_distance_py(s1, s2, window: int, penalty: float) -> float:
    """Pure-Python statement of SURVEY.md App. A.1 (full matrix, no rolling
    rows) — a second, independently written form used to cross-check the C."""
    s1, s2 = np.asarray(s1, dtype=np.float64), np.asarray(s2, dtype=np.float64)
    r, c = len(s1), len(s2)
    if window is None or window <= 0:
        window = max(r, c)
    p2 = np.float64(penalty) * np.float64(penalty)
    D = np.full((r + 1, c + 1), np.inf)
    D[0, 0] = 0.0
    for i in range(r):
        for j in range(max(0, i - max(0, r - c) - window + 1), min(c, i + max(0, c - r) + window)):
            d = (s1[i] - s2[j]) * (s1[i] - s2[j])
            m = D[i, j]
            t = D[i, j + 1] + p2
            if t < m:
                m = t
            t = D[i + 1, j] + p2
            if t < m:
                m = t
            D[i + 1, j + 1] = d + m
    return float(np.sqrt(D[r, c]))


def dtw_matrix(X, Y, window: int, penalty: float) -> np.ndarray:
    """float64 [nX, nY] DTW distances (before the reference's float32 cast)."""
    X, Y = _c64(np.atleast_2d(X)), _c64(np.atleast_2d(Y))
    assert X.shape[1] == Y.shape[1]
    out = np.empty((X.shape[0], Y.shape[0]), dtype=np.float64)
    lib().wdx_oracle_dtw_matrix(_d(X), X.shape[0], _d(Y), Y.shape[0], X.shape[1], int(window or 0),
                                float(penalty or 0.0), _d(out))
    return out


def distance_matrix_to(X, Y, window=None, penalty=None, **_ignored) -> np.ndarray:
    """`warpdemux.parallel_distances.distance_matrix_to` (n_jobs=1 branch,
    parallel_distances.py:58-67): float64 DTW then `.astype(np.float32)`."""
    return dtw_matrix(X, Y, window, penalty).astype(np.float32)


# --------------------------------------------------------------------------
# DTW_SVM.predict
# --------------------------------------------------------------------------
def pdist_kernel(pdist: np.ndarray, gamma: float = 1, pwr_dist: int = 1) -> np.ndarray:
    """dtw_svm.py:21-22, verbatim arithmetic (float32 in, float32 out)."""
    return np.exp(-gamma * np.power(pdist, pwr_dist))


def svc_predict_proba(K: np.ndarray, m) -> Tuple[np.ndarray, np.ndarray]:
    """libsvm predict_probability over a float32 kernel matrix; returns
    (prob [n,k], dec [n,k(k-1)/2])."""
    K = np.ascontiguousarray(K, dtype=np.float32)
    n = K.shape[0]
    assert K.shape[1] == m.n_sv
    prob = np.empty((n, m.k), dtype=np.float64)
    dec = np.empty((n, m.n_pairs), dtype=np.float64)
    rc = lib().wdx_oracle_svc_predict_proba(
        K.ctypes.data_as(_fp), n, m.n_sv, m.k, m.n_sv_class.ctypes.data_as(_ip), _d(m.dual_coef), _d(m.rho),
        _d(m.probA), _d(m.probB), _d(dec), _d(prob))
    if rc != 0:
        raise ValueError("oracle: unsupported class count")
    return prob, dec


def confidence_margin(npa: np.ndarray) -> np.ndarray:
    """models/utils.py:19-22."""
    s = np.sort(npa, axis=1)[:, ::-1]
    return s[:, 0] - s[:, 1]


def process_probs(y_prob: np.ndarray, m) -> Tuple[np.ndarray, np.ndarray]:
    """models/utils.py:45-61 with the model's label_mapper / thresholds."""
    pred_idx = np.argmax(y_prob, axis=1)
    pred = m.label_map[pred_idx].copy()
    conf = confidence_margin(y_prob)
    pred[conf < m.thresholds[pred_idx]] = -1
    return pred, conf


def predict(m, X: np.ndarray):
    """DTW_SVM.predict (dtw_svm.py:54-98) -> (y_pred int64[n], y_prob f64[n,k],
    conf f64[n], D float32[n,n_sv])."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    if X.shape[1] != m.sv.shape[1]:
        raise ValueError("X must have the same number of columns as the training data " f" ({m.sv.shape[1]}).")
    D = distance_matrix_to(X, m.sv, window=m.window, penalty=m.penalty)
    K = pdist_kernel(D, gamma=m.gamma, pwr_dist=m.pwr_dist)
    prob, _ = svc_predict_proba(K, m)
    pred, conf = process_probs(prob, m)
    return pred, prob, conf, D


def predict_c(m, X: np.ndarray):
    """Whole path inside one C call (timed CPU arm). Same arithmetic except
    float32 exp is libm expf instead of numpy's SIMD exp."""
    X = _c64(np.atleast_2d(X))
    n = X.shape[0]
    pred = np.empty(n, dtype=np.int64)
    conf = np.empty(n, dtype=np.float64)
    prob = np.empty((n, m.k), dtype=np.float64)
    rc = lib().wdx_oracle_predict(
        _d(X), n, _d(m.sv), m.n_sv, m.L, int(m.window), float(m.penalty), float(m.gamma), int(m.pwr_dist), m.k,
        m.n_sv_class.ctypes.data_as(_ip), _d(m.dual_coef), _d(m.rho), _d(m.probA), _d(m.probB),
        m.label_map.ctypes.data_as(_lp), _d(m.thresholds), pred.ctypes.data_as(_lp), _d(conf), _d(prob))
    if rc != 0:
        raise ValueError("oracle: unsupported class count")
    return pred, prob, conf


def predict_threaded(m, X: np.ndarray, threads: int, minibatch: int = 1000):
    """Production-style CPU parallelism (file_proc.py:1197-1245): minibatches
    of 1000 reads over a pool of single-threaded workers."""
    from concurrent.futures import ThreadPoolExecutor

    X = _c64(np.atleast_2d(X))
    chunks = [X[i:i + minibatch] for i in range(0, X.shape[0], minibatch)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda c: predict_c(m, c), chunks))
    return (np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]),
            np.concatenate([r[2] for r in res]))


# --------------------------------------------------------------------------
# Fingerprint extraction (sig_proc.py:394-605, non-consensus path)
# --------------------------------------------------------------------------
def windowed_t_test(x: np.ndarray, w: int) -> np.ndarray:
    """segmentation.py:32-45 -> _c_segmentation.pyx:124-161."""
    x = _c64(x)
    nc = x.size - 2 * w
    if nc <= 0:
        return np.zeros(0, dtype=np.float64)
    out = np.empty(nc, dtype=np.float64)
    lib().wdx_oracle_windowed_t_test(_d(x), x.size, int(w), _d(out))
    return out


def new_means(x: np.ndarray, segs: np.ndarray) -> np.ndarray:
    """segmentation.py:48-74 -> _c_segmentation.pyx:41-53."""
    x = _c64(x)
    segs = np.ascontiguousarray(segs, dtype=np.int64)
    out = np.empty(segs.size - 1, dtype=np.float64)
    lib().wdx_oracle_new_means(_d(x), segs.ctypes.data_as(_lp), segs.size - 1, _d(out))
    return out


FP_OK = 0
FP_FAIL_SEGMENTATION = 1  # "event segmentation failed" (sig_proc.py:537-544)
FP_FAIL_DETECT = 2        # detect_results.success == False (sig_proc.py:400-407)
FP_FAIL_NORMALIZE = 3     # "segment normalization failed" (sig_proc.py:553-560)


def fingerprint(signal: np.ndarray, adapter_start: int, adapter_end: int, *, padding: int = 100,
                outlier_thresh: float = 5.0, min_obs_per_base: int = 6, running_stat_width: int = 12,
                num_events: int = 110, barcode_num_events: int = 25, mutate: bool = False):
    """`detect_results_to_fpt` (sig_proc.py:394-605) for the configuration every
    shipped DTW-SVM model uses (rna004_130bps@v1.0.toml: sig_extract
    normalization "none", segmentation normalization "mean",
    accept_less_cpts false, no consensus refinement).

    Returns (status, fpt float64[25], dwell int64[25], stats dict).
    numpy 2 scalar semantics (NEP 50) apply to the clip bounds: float32.
    """
    from scipy.signal import find_peaks

    signal = np.asarray(signal)
    start = max(0, adapter_start - padding)                      # sig_proc.py:382-391
    stop = min(signal.size, adapter_end + padding)
    sig = signal[start:stop]
    if not mutate:
        sig = sig.copy()
    med = np.nanmedian(sig)                                      # :421
    mad = np.nanmedian(np.abs(sig - med))                        # :422
    np.clip(sig, med - outlier_thresh * mad, med + outlier_thresh * mad, out=sig)   # :426-431
    n = sig.size
    m_obs = min(min_obs_per_base, round(n / num_events / 2))     # :526-529
    w = min(running_stat_width, round(n / num_events))           # :530-533
    empty = (np.full(barcode_num_events, np.nan), np.zeros(barcode_num_events, dtype=np.int64), {})
    scores = windowed_t_test(sig, w)                             # :225-228
    try:
        peaks, _ = find_peaks(scores, distance=m_obs)            # :183
    except ValueError:                                           # distance < 1 (very short adapter)
        return (FP_FAIL_SEGMENTATION,) + empty
    if peaks.size < num_events:                                  # :185-186
        return (FP_FAIL_SEGMENTATION,) + empty
    cpts = peaks[np.argsort(scores[peaks])[-num_events:]] + w    # :188
    cpts.sort()
    if cpts[0] != 0:
        cpts = np.insert(cpts, 0, 0)
    if cpts[-1] != n:
        cpts = np.append(cpts, n)
    dwell = cpts[1:] - cpts[:-1]                                 # :241
    ev = new_means(sig, cpts)                                    # :242
    if np.isnan(ev).any():                                       # normalize(..., accept_nan=False) :104-107
        return (FP_FAIL_NORMALIZE,) + empty
    norm = (ev - np.mean(ev, axis=-1, keepdims=True)) / np.std(ev, axis=-1, keepdims=True)   # :99-111
    dmed = float(np.median(dwell))
    stats = dict(
        adapter_dt_med=dmed,
        adapter_dt_mad=float(np.median(np.abs(dwell - dmed))),
        adapter_event_mean=float(ev.mean()),
        adapter_event_std=float(ev.std()),
        adapter_event_med=float(np.median(ev)),
        adapter_event_mad=float(np.median(np.abs(ev - np.median(ev)))),
    )
    keep = min(barcode_num_events, norm.size)
    return FP_OK, norm[-keep:], dwell[-keep:].astype(np.int64), stats
