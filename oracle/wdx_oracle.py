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
        L.wdx_oracle_warping_paths.restype = None
        L.wdx_oracle_warping_paths.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, C.c_int, C.c_int, _dp]
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


def select_by_peak_distance_stable(peaks: np.ndarray, priority: np.ndarray, distance: int) -> np.ndarray:
    """scipy.signal._peak_finding_utils._select_by_peak_distance with a STABLE argsort: of two equally high peaks closer
    than `distance` the later one stays.  scipy sorts with numpy's default (unstable) argsort, so which of the two stays
    is not defined by the reference — it depends on the numpy build and the CPU's SIMD level.  Returns the keep mask."""
    n = peaks.size
    keep = np.ones(n, dtype=bool)
    order = np.argsort(priority, kind="stable")
    for i in range(n - 1, -1, -1):
        j = order[i]
        if not keep[j]:
            continue
        k = j - 1
        while 0 <= k and peaks[j] - peaks[k] < distance:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < distance:
            keep[k] = False
            k += 1
    return keep


def fingerprint_has_ties(signal: np.ndarray, adapter_start: int, adapter_end: int, *, padding: int = 100,
                         outlier_thresh: float = 5.0, min_obs_per_base: int = 6, running_stat_width: int = 12,
                         num_events: int = 110, **_ignored) -> bool:
    """True when the reference's change points depend on how numpy's unstable argsort orders EQUAL t-test scores:
    two equally high local maxima closer than the peak distance, or a tie at the num_events-th highest kept peak."""
    from scipy.signal import find_peaks

    signal = np.asarray(signal)
    sig = signal[max(0, adapter_start - padding):min(signal.size, adapter_end + padding)].copy()
    med = np.nanmedian(sig)
    mad = np.nanmedian(np.abs(sig - med))
    np.clip(sig, med - outlier_thresh * mad, med + outlier_thresh * mad, out=sig)
    n = sig.size
    m_obs = min(min_obs_per_base, round(n / num_events / 2))
    w = min(running_stat_width, round(n / num_events))
    if m_obs < 1 or w < 1 or n - 2 * w < 3:
        return False
    scores = windowed_t_test(sig, w)
    pk, _ = find_peaks(scores)
    for a in range(pk.size - 1):
        b = a + 1
        while b < pk.size and pk[b] - pk[a] < m_obs:
            if scores[pk[a]] == scores[pk[b]]:
                return True
            b += 1
    keep = select_by_peak_distance_stable(pk, scores[pk], m_obs)
    h = np.sort(scores[pk[keep]])
    return bool(h.size > num_events and h[-num_events] == h[-num_events - 1])


def fingerprint(signal: np.ndarray, adapter_start: int, adapter_end: int, *, padding: int = 100,
                outlier_thresh: float = 5.0, min_obs_per_base: int = 6, running_stat_width: int = 12,
                num_events: int = 110, barcode_num_events: int = 25, mutate: bool = False, stable_ties: bool = False,
                numpy1_promotion: bool = False):
    """`detect_results_to_fpt` (sig_proc.py:394-605) for the configuration every
    shipped DTW-SVM model uses (rna004_130bps@v1.0.toml: sig_extract
    normalization "none", segmentation normalization "mean",
    accept_less_cpts false, no consensus refinement).

    Returns (status, fpt float64[25], dwell int64[25], stats dict).
    numpy 2 scalar semantics (NEP 50) apply to the clip bounds: float32.
    stable_ties: equal t-test scores are ordered by a stable sort (the later peak wins) instead of by numpy's default
    argsort, whose order of equal keys is unspecified (introsort / AVX-512 / AVX2 sorting networks depending on the numpy
    build and the CPU) — the one point where the reference's own result is machine-dependent.
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
    if numpy1_promotion:   # numpy < 2 (the reference pins 1.26.4): np.float32 scalar * Python float -> float64; np.clip casts once
        lo = np.float32(np.float64(med) - outlier_thresh * np.float64(mad))
        hi = np.float32(np.float64(med) + outlier_thresh * np.float64(mad))
        np.clip(sig, lo, hi, out=sig)
    else:
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
    if stable_ties:      # the tie rule of the CUDA kernels (see select_by_peak_distance_stable)
        peaks, _ = find_peaks(scores)
        peaks = peaks[select_by_peak_distance_stable(peaks, scores[peaks], m_obs)]
    if peaks.size < num_events:                                  # :185-186
        return (FP_FAIL_SEGMENTATION,) + empty
    cpts = peaks[np.argsort(scores[peaks], kind="stable" if stable_ties else None)[-num_events:]] + w    # :188
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


# --------------------------------------------------------------------------
# Consensus-guided barcode refinement (tRNA configs; sig_proc.py:257-378, 451-521)
# PARITY UNPINNED for the dtaidistance pieces (warping_paths, SubsequenceAlignment,
# best_path): restated from the published 2.3.13 algorithm, no upstream source,
# wheel or reference vector is available here.
# --------------------------------------------------------------------------
FP_FAIL_TOO_LONG = 4
FP_FAIL_CONSENSUS = 5     # "consensus query outlier" (sig_proc.py:500-521)


def warping_paths(s1: np.ndarray, s2: np.ndarray, penalty: float, psi) -> Tuple[float, np.ndarray]:
    """dtaidistance `dtw.warping_paths_fast(s1, s2, penalty=, psi=(b1, e1, b2, e2), compact=False,
    psi_neg=False)` for e1 = e2 = 0 (the reference's call, sig_proc.py:298-305): (distance, paths) with
    paths the sqrt'ed (len(s1)+1) x (len(s2)+1) cumulative-cost matrix."""
    s1 = _c64(s1)
    s2 = _c64(s2)
    psi_1b, psi_1e, psi_2b, psi_2e = (int(v) for v in psi)
    if psi_1e or psi_2e:
        raise NotImplementedError("end relaxation is not used by the reference")
    out = np.empty((s1.size + 1, s2.size + 1), dtype=np.float64)
    lib().wdx_oracle_warping_paths(_d(s1), s1.size, _d(s2), s2.size, float(penalty), psi_1b, psi_2b, _d(out))
    return float(out[s1.size, s2.size]), out


def best_path(paths: np.ndarray, row: Optional[int] = None, col: Optional[int] = None):
    """dtaidistance `dtw.best_path(paths, row, col)`: walk back from (row, col) to the border, at
    every step to the FIRST minimum of (diagonal, up, left); returns [(i-1, j-1), ...] in forward
    order without the border cell."""
    i = paths.shape[0] - 1 if row is None else row
    j = paths.shape[1] - 1 if col is None else col
    p = []
    if paths[i, j] != -1:
        p.append((i - 1, j - 1))
    while i > 0 and j > 0:
        c, vmin = 0, float("inf")
        for q, v in enumerate((paths[i - 1, j - 1], paths[i - 1, j], paths[i, j - 1])):
            if v < vmin:
                c, vmin = q, v
        if c == 0:
            i, j = i - 1, j - 1
        elif c == 1:
            i = i - 1
        else:
            j = j - 1
        if paths[i, j] != -1:
            p.append((i - 1, j - 1))
    p.pop()
    p.reverse()
    return p


def subsequence_best_match(query: np.ndarray, series: np.ndarray, penalty: float, psi) -> Tuple[int, int]:
    """`_get_subseq_match` (sig_proc.py:288-312): SubsequenceAlignment with externally computed paths,
    `_compute_matching` (last row / len(query), trimmed to len(series)), `best_match().segment`
    = [start of best_path(paths, col=idx+1), idx]."""
    _, paths = warping_paths(query, series, penalty, psi)
    matching = paths[-1, :]
    if matching.size > series.size:
        matching = matching[-series.size:]
    matching = np.array(matching) / len(query)
    best_idx = int(np.argmin(matching))
    path = best_path(paths, col=best_idx + 1)
    return int(path[0][1]), best_idx


def _cpts_from_scores(scores, num_events, min_obs, w):
    """discrepenacy_curve_to_cpts (sig_proc.py:176-198), accept_less_cpts=False; None on failure."""
    from scipy.signal import find_peaks

    try:
        peaks, _ = find_peaks(scores, distance=min_obs)
    except ValueError:
        return None
    if peaks.size < num_events:
        return None
    cpts = peaks[np.argsort(scores[peaks])[-num_events:]] + w
    cpts.sort()
    signal_len = scores.size + 2 * w
    if cpts[0] != 0:
        cpts = np.insert(cpts, 0, 0)
    if cpts[-1] != signal_len:
        cpts = np.append(cpts, signal_len)
    return cpts


def fingerprint_consensus(signal: np.ndarray, adapter_start: int, adapter_end: int, consensus: np.ndarray, *,
                          padding: int = 100, outlier_thresh: float = 5.0, min_obs_per_base: int = 9,
                          running_stat_width: int = 18, num_events: int = 120, barcode_num_events=(25, 25),
                          penalty: float = 1.5, psi=(5, 0, 40, 0), ub_start: int = 18, lb_end: int = 69,
                          ub_end: int = 97, mutate: bool = False):
    """`detect_results_to_fpt` with `consensus_refinement = true` (sig_proc.py:451-521 ->
    segment_signal_with_consensus_guided_barcode_refinement :257-378), defaults of
    rna004_130bps@v1.0_tRNA.toml:13-29 (sig_extract normalization "none", segmentation and
    consensus normalization "mean", refinement_optimal_cpts false).

    Returns (status, fpt[retain], dwell[retain], stats dict, (seg_query_start, seg_query_end,
    sig_barcode_start)).  Two inputs on which the reference itself does not return are given a
    status instead: NaN padding inside the slice (normalize() raises inside _get_subseq_match ->
    FP_FAIL_NORMALIZE) and adapters so short that the adapter window is narrower than
    running_stat_width (compute_base_means is handed change points beyond the signal end, an
    out-of-bounds read in the reference -> FP_FAIL_SEGMENTATION).
    """
    seg_events, retain = int(barcode_num_events[0]), int(barcode_num_events[1])
    signal = np.asarray(signal)
    start = max(0, adapter_start - padding)
    stop = min(signal.size, adapter_end + padding)
    sig = signal[start:stop]
    if not mutate:
        sig = sig.copy()
    med = np.nanmedian(sig)
    mad = np.nanmedian(np.abs(sig - med))
    np.clip(sig, med - outlier_thresh * mad, med + outlier_thresh * mad, out=sig)
    n = sig.size
    empty = (np.full(retain, np.nan), np.zeros(retain, dtype=np.int64), {}, (0, 0, 0))
    m_obs = min(min_obs_per_base, int(round(n / num_events / 2)))          # :315-318
    w = min(running_stat_width, int(round(n / num_events)))                # :319-322
    scores = windowed_t_test(sig, w)
    cpts = _cpts_from_scores(scores, num_events, m_obs, w)
    if cpts is None:
        return (FP_FAIL_SEGMENTATION,) + empty
    a_dwell = cpts[1:] - cpts[:-1]
    a_ev = new_means(sig, cpts)
    if np.isnan(a_ev).any():
        return (FP_FAIL_NORMALIZE,) + empty
    if w != running_stat_width:
        return (FP_FAIL_SEGMENTATION,) + empty
    norm_series = (a_ev - np.mean(a_ev, axis=-1, keepdims=True)) / np.std(a_ev, axis=-1, keepdims=True)
    q_start, q_end = subsequence_best_match(np.asarray(consensus, dtype=np.float64), norm_series, penalty, psi)
    sig_bc_start = int(np.sum(a_dwell[:q_end]))                            # :334
    bc_scores = scores[sig_bc_start:]                                      # :336
    bc_cpts = _cpts_from_scores(bc_scores, seg_events, min_obs_per_base, running_stat_width)   # :359-365
    if bc_cpts is None:
        return (FP_FAIL_SEGMENTATION,) + empty
    bc_dwell = bc_cpts[1:] - bc_cpts[:-1]
    bc_ev = new_means(sig[sig_bc_start:], bc_cpts)                         # :371
    norm = ((bc_ev[:, None] - np.mean(a_ev)) / np.std(a_ev)).squeeze(-1)   # normalize_wrt :139-168
    dmed = float(np.median(a_dwell))
    stats = dict(
        adapter_dt_med=dmed,
        adapter_dt_mad=float(np.median(np.abs(a_dwell - dmed))),
        adapter_event_mean=float(a_ev.mean()),
        adapter_event_std=float(a_ev.std()),
        adapter_event_med=float(np.median(a_ev)),
        adapter_event_mad=float(np.median(np.abs(a_ev - np.median(a_ev)))),
    )
    cons = (q_start, q_end, sig_bc_start)
    if q_start > ub_start or q_end < lb_end or q_end > ub_end:             # :500-521
        return FP_FAIL_CONSENSUS, empty[0], empty[1], stats, cons
    keep = min(retain, norm.size)
    return FP_OK, norm[-keep:], bc_dwell[-keep:].astype(np.int64), stats, cons
