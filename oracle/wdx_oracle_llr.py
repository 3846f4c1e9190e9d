"""CPU restatement of the LLR fallback of the adapter / poly(A) boundary detection — the branch of
`combined_detect_cnn` that re-detects the reads whose CNN boundaries fail validation.
TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by golden generators; never by the
product (warpdemux_b200/).

Follows (warpdemux/adapted/adapted/detect/):
    combined_detect_cnn, fallback branch         combined.py:222-296
    detect_llr_on_downscaled_signal              combined.py:39-129
    downscale_single_read_excl_nan               combined.py:132-142, downscale.py:4-41
    normalize_signal / med_mad / clip_signal     normalize.py:15-63        (float32 in, float32 out)
    calc_adapter_trace -> c_llr_trace / _gains   llr.py:243-334, _c_llr.pyx:24-38, 66-88, 184-230   (C: wdx_oracle_llr_gains)
    LLRTrace._trace_start_end                    llr.py:113-121
    find_peaks_in_trace, adapter_end_from_trace  llr.py:181-240
    correct_for_plateau, correct_for_split_peak  llr.py:124-178
    detect_full_polya_trace_peak_with_spike      llr.py:385-455
scipy's `find_peaks` (prominence, width, rel_height, distance) and `linregress` are the real scipy functions here, as
in the reference; the CUDA kernel restates them.  The gains are pinned on the reference's own compiled Cython
(oracle/_ref/ref_c_llr*.so, tests/test_oracle_llr.py), the whole branch on the reference's results for all 4000 reads
of test_data/demux (tests/golden/real4000_rna004_WDX4.npz: 28 hail-mary and 193 LLR re-detections).

`log` is the C library's (the reference's Cython calls libc `log`); which libm variant runs is machine-dependent
(glibc selects an FMA build at load time), so the float traces are reproducible to ~1 ulp only — the integer
boundaries are what is compared.
"""
from __future__ import annotations

import dataclasses
import warnings

import numpy as np

from . import wdx_oracle as _o
from . import wdx_oracle_validate as _v

# fail codes on top of wdx_oracle_validate's (exceptions the reference's try/except turns into fail_reason strings)
MAD_ZERO = 10            # ValueError("MAD normalization failed: scale is 0")   normalize.py:55-58
LLR_ERROR = 11           # any other exception inside the fallback branch (degenerate traces)
FAIL_REASON = dict(_v.FAIL_REASON)
FAIL_REASON[MAD_ZERO] = "MAD normalization failed: scale is 0"
FAIL_REASON[LLR_ERROR] = "LLR detection failed"

PATH_CNN, PATH_HAIL_MARY, PATH_LLR = 0, 1, 2


@dataclasses.dataclass
class LLRConfig:
    """SigProcConfig fields the fallback reads (WarpDemuX rna004_130bps@v1.0.toml on top of adapted @v0.2.4)."""
    max_obs_trace: int = 10000
    min_obs_adapter: int = 1000
    max_obs_adapter: int = 6500
    downscale_factor: int = 10
    sig_norm_outlier_thresh: float = 5.0
    adapter_peak_prominence: float = 1.0
    adapter_peak_rel_height: float = 1.0
    adapter_peak_width: int = 1000
    fallback_to_llr: bool = True
    fallback_to_llr_short_reads: bool = True


def _lib():
    import ctypes as C

    L = _o.lib()
    if not getattr(L, "_llr_ready", False):
        L.wdx_oracle_cumsums.restype = None
        L.wdx_oracle_cumsums.argtypes = [_o._dp, C.c_int64, _o._dp, _o._dp]
        L.wdx_oracle_llr_gains.restype = None
        L.wdx_oracle_llr_gains.argtypes = [_o._dp, _o._dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _o._dp]
        L._llr_ready = True
    return L


def cumsums(x: np.ndarray):
    """c = np.cumsum(x), c2 = np.cumsum(x * x) (float64, sequential; _c_llr.pyx:214-215)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    c, c2 = np.empty_like(x), np.empty_like(x)
    if x.size:
        _lib().wdx_oracle_cumsums(_o._d(x), x.size, _o._d(c), _o._d(c2))
    return c, c2


def gains(c: np.ndarray, c2: np.ndarray, start: int, end: int, offset_head: int, offset_tail: int) -> np.ndarray:
    """_gains(start, end, c, c2, offset_head, offset_tail, stride=1) (_c_llr.pyx:66-88)."""
    g = np.zeros_like(c)
    if c.size:
        if end - 1 >= c.size or end < start or start < 0:
            raise IndexError("llr gains: segment out of range")
        with np.errstate(all="ignore"):
            _lib().wdx_oracle_llr_gains(_o._d(c), _o._d(c2), c.size, int(start), int(end), int(offset_head), int(offset_tail), _o._d(g))
    return g


def normalize_signal(signal: np.ndarray, outlier_thresh: float) -> np.ndarray:
    """normalize_signal(signal, outlier_thresh, with_nan=True) (normalize.py:32-63): float32 result for float32 rows."""
    if len(signal) == 0:
        return np.array([], dtype=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        med = float(np.nanmedian(signal))
        mad = float(np.nanmedian(np.abs(signal - med)))
    if mad == 0:
        raise ValueError("MAD normalization failed: scale is 0")
    norm = np.clip(signal, med - (mad * outlier_thresh), med + (mad * outlier_thresh))
    return (norm - med) / mad


def downscale_excl_nan(signal: np.ndarray, factor: int) -> np.ndarray:
    """downscale_single_read_excl_nan (combined.py:132-142): zero-padded block means, NaN blocks (counted, assumed to be
    at the end) dropped."""
    x = signal.reshape(1, -1)
    n = x.shape[1]
    if n % factor:
        x = np.pad(x, ((0, 0), (0, factor - n % factor)), mode="constant")
    ds = x.reshape(1, -1, factor).mean(axis=2).ravel()
    return ds[: ds.size - int(np.isnan(ds).sum())]


def trace_start_end(sig: np.ndarray):
    """LLRTrace._trace_start_end (llr.py:113-121)."""
    start = int(np.argmin(sig <= 0))
    end = int(sig.size - np.argmin(sig[::-1] <= 0) - 1)
    return start, end


def correct_for_plateau(trace_sig, peak, s=10, t=0.9, window=500):
    tr = trace_sig[peak:min(peak + window, trace_sig.size)]
    plateau_end = -1
    changes = np.diff(tr)
    n = len(changes)
    for i in range(n - s, -1, -1):
        if (changes[i:i + (s - 1)] >= 0).all() and tr[i + (s - 1)] > t * tr[0]:
            plateau_end = i + (s - 1)
            break
    return peak + plateau_end if plateau_end > 0 else peak


def correct_for_split_peak(trace_sig, peak, s=10, t=0.9, window=500, prominence=1.0):
    from scipy.signal import find_peaks

    peaks, _ = find_peaks(trace_sig[peak:min(peak + window, trace_sig.size)], width=s, prominence=prominence)
    if peaks.size > 0 and trace_sig[peaks[0] + peak] >= t * trace_sig[peak]:
        return int(peaks[0] + peak)
    return int(peak)


def adapter_end_from_trace(sig: np.ndarray, width: int, prominence: float, rel_height: float):
    """adapter_end_from_trace(trace, ..., fix_plateau=True, correct_for_split_peaks=True) (llr.py:204-240); only the
    first candidate is used by the caller.  Returns the candidate list."""
    from scipy.signal import find_peaks

    start, end = trace_start_end(sig)
    clip = sig[start:end]
    peaks, _ = find_peaks(clip, width=width, prominence=prominence * np.nanstd(clip), rel_height=rel_height)
    peaks = peaks + start
    peaks = [correct_for_plateau(sig, int(p)) for p in peaks]
    peaks = [correct_for_split_peak(sig, int(p)) for p in peaks]
    return peaks


def polya_trace_peak_with_spike(trace: np.ndarray, min_peak_distance=10, prominence_threshold=1.0, min_width=10,
                                threshold_prominence_ratio=0.5, threshold_r_squared=0.99) -> int:
    """detect_full_polya_trace_peak_with_spike (llr.py:385-455)."""
    from scipy.signal import find_peaks
    from scipy.stats import linregress

    peaks, _ = find_peaks(np.nan_to_num(trace, nan=0), distance=min_peak_distance, prominence=prominence_threshold,
                          width=min_width, rel_height=0.5)
    if len(peaks) == 0:
        return 0
    if len(peaks) == 1:
        return int(peaks[0])
    h = trace[peaks]
    if h[1] > h[0]:
        return int(peaks[1])
    if h[1] < h[0] * threshold_prominence_ratio:
        return int(peaks[0])
    idx_min = trace[peaks[0]:peaks[1]].argmin() + peaks[0]
    x2 = np.arange(idx_min, peaks[1])
    r = linregress(x2, trace[x2])[2]
    return int(peaks[1]) if r ** 2 >= threshold_r_squared else 0


def detect_llr_on_downscaled_signal(s: np.ndarray, cfg: LLRConfig):
    """combined.py:39-129 -> (adapter_end, polya_end) in samples (0 = none)."""
    f = cfg.downscale_factor
    x = np.ascontiguousarray(s, dtype=np.float64)
    n = x.size
    c, c2 = cumsums(x)
    tr = gains(c, c2, 0, n - 1, 1 + cfg.min_obs_adapter // f, 1)
    adapter_end = polya_end = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cands = adapter_end_from_trace(tr, cfg.adapter_peak_width // f, cfg.adapter_peak_prominence, cfg.adapter_peak_rel_height)
        if len(cands) > 0 and cands[0] > 0:
            ae = int(cands[0])
            adapter_end = ae * f
            tr2 = gains(c, c2, ae, n - 1, 1, 1)
            pe = polya_trace_peak_with_spike(tr2)
            if pe > 0:
                polya_end = pe * f
    return adapter_end, polya_end


def hail_mary_polya(norm_signal: np.ndarray, adapter_end: int, polya_end: int, cfg: LLRConfig) -> int:
    """combined.py:242-266: poly(A) end re-detected on the CNN's [adapter_end, polya_end) stretch; 0 = none."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = downscale_excl_nan(norm_signal[int(adapter_end):int(polya_end)], cfg.downscale_factor)
        x = np.ascontiguousarray(s, dtype=np.float64)
        c, c2 = cumsums(x)
        tr = gains(c, c2, 0, x.size - 1, 5, 5)
        pe = polya_trace_peak_with_spike(tr)
    return int(pe * cfg.downscale_factor + adapter_end) if pe > 0 else 0


def detect_one(row: np.ndarray, full_signal_len: int, cnn_pred, cfg: LLRConfig, vcfg: "_v.ValidateConfig"):
    """One read through combined_detect_cnn after the CNN (combined.py:209-296): validation of the CNN boundaries, then
    the hail-mary and LLR re-detections.  cnn_pred = [adapter_end, polya_end candidates...].
    Returns (validate_one-style dict, path, info) with info = dict(hm_tried, hm_polya, llr_tried, llr_adapter_end,
    llr_polya_end, primary) — `primary` is the method whose boundaries the returned result validated."""
    info = dict(hm_tried=False, hm_polya=0, llr_tried=False, llr_adapter_end=0, llr_polya_end=0, primary="cnn")
    ae, topk = int(cnn_pred[0]), np.asarray(cnn_pred[1:], dtype=np.int64)
    res = _v.validate_one(row, full_signal_len, ae, topk, vcfg)
    path = PATH_CNN
    if res["success"] or res["code"] == _v.HAS_NAN:
        return res, path, info
    pe = int(topk[0]) if topk.size else 0
    try:
        sig = np.asarray(row)
        norm = normalize_signal(sig[: min(cfg.max_obs_trace, full_signal_len)], cfg.sig_norm_outlier_thresh)
        if (ae > 0 and pe > 0 and pe - ae > 1000 and full_signal_len < 2 * cfg.max_obs_adapter
                and cfg.fallback_to_llr_short_reads):
            new_pe = hail_mary_polya(norm, ae, pe, cfg)
            info["hm_tried"] = True
            if new_pe > 0:
                info["hm_polya"] = new_pe
                res = _v.validate_one(row, full_signal_len, ae, np.array([new_pe]), vcfg)
                info["primary"] = "llr"
                info["llr_adapter_end"], info["llr_polya_end"] = ae, new_pe
                if res["success"]:
                    path = PATH_HAIL_MARY
        if not res["success"] and cfg.fallback_to_llr:
            s = downscale_excl_nan(norm[: min(cfg.max_obs_trace, full_signal_len)], cfg.downscale_factor)
            l_ae, l_pe = detect_llr_on_downscaled_signal(s, cfg)
            info["llr_tried"] = True
            r2 = _v.validate_one(row, full_signal_len, l_ae, np.array([l_pe]) if l_pe > 0 else np.zeros(0, np.int64), vcfg)
            if r2["success"]:
                res, path = r2, PATH_LLR
                info["primary"] = "llr"
                info["llr_adapter_end"], info["llr_polya_end"] = l_ae, l_pe
    except ValueError as e:
        code = MAD_ZERO if "MAD normalization" in str(e) else LLR_ERROR
        res = dict(res, success=False, code=code, checks=0)
    except Exception:  # noqa: BLE001
        res = dict(res, success=False, code=LLR_ERROR, checks=0)
    return res, path, info
