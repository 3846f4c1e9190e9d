"""CPU restatement of the boundary VALIDATION step that sits between the boundary CNN and the
fingerprint stage.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by the golden generator; never by the product (warpdemux_b200/).

Follows, for `mvs_detect_overwrite = false` (every shipped config):
    validate_boundaries                 warpdemux/adapted/adapted/detect/combined.py:409-683
    find_open_pores                     adapted/detect/anomalies.py:16-35
    real_range_check                    adapted/detect/real_range.py:34-63
    mean_var_shift_polyA_check          adapted/detect/mvs.py:42-159
    in_range                            adapted/detect/utils.py:16-26
The third-party `bottleneck` package (move_mean / move_var, mvs.py:93-107) is absent from this image;
its documented semantics are restated here the same way as in oracle/shim/bottleneck (trailing window,
min_count = window, ddof = 0, float32 in -> float32 out, float64 accumulation) -> **parity unpinned**
for those two statistics; everything else is pinned on the reference's own validate_boundaries
(tests/golden/validate_rna004.npz, oracle/make_golden_validate.py).
The partition statistics of DetectResults (adapted/partition/signal_partitions.py:65-96: start, len, mean, std, med,
mad of the adapter / poly(A) / preloaded-RNA partitions, reporting only) are restated in `partitions()`.
"""
from __future__ import annotations

import dataclasses
import math
import warnings

import numpy as np

# fail codes (the reference's fail_reason strings, combined.py)
OK = 0
NO_ADAPTER = 1          # "No adapter detected (primary)"
ADAPTER_MAD = 2         # "adapter MAD check failed"
OPEN_PORE = 3           # "Open pore too close to boundary"
REAL_RANGE = 4          # "Real signal check failed"
NO_POLYA = 5            # "No polya detected (primary)"
MVS_NO_SIGNAL = 6       # "MVS polya check failed: not enough signal"
MVS_CHECKS = 7          # "MVS polya check failed: <failed checks>"
MED_SHIFT = 8           # "Median shift check failed"
HAS_NAN = 9             # ValueError("Validate boundaries failed: Signal contains nan values") -> caught by the caller

FAIL_REASON = {
    OK: None,
    NO_ADAPTER: "No adapter detected (primary)",
    ADAPTER_MAD: "adapter MAD check failed",
    OPEN_PORE: "Open pore too close to boundary",
    REAL_RANGE: "Real signal check failed",
    NO_POLYA: "No polya detected (primary)",
    MVS_NO_SIGNAL: "MVS polya check failed: not enough signal",
    MVS_CHECKS: "MVS polya check failed: ",
    MED_SHIFT: "Median shift check failed",
    HAS_NAN: "Validate boundaries failed: Signal contains nan values",
}
CHECK_NAMES = ("mean", "var", "med", "range", "shift")

# columns of the per-read value vector (NaN = the reference's None)
V_ADAPTER_MED, V_ADAPTER_MAD, V_REAL_MEAN_START, V_REAL_MEAN_END, V_REAL_LOCAL_RANGE = 0, 1, 2, 3, 4
V_MVS_MEAN, V_MVS_VAR, V_MVS_POLYA_MED, V_MVS_POLYA_LOCAL_RANGE, V_MVS_MED_SHIFT, V_ADAPTER_RNA_MED_SHIFT = 5, 6, 7, 8, 9, 10
N_VALS = 12

INF = float("inf")


@dataclasses.dataclass
class ValidateConfig:
    """The fields of SigProcConfig that validate_boundaries reads (defaults: rna004_130bps@v0.2.4.toml)."""
    min_obs_adapter: int = 1000
    # [real_range]
    detect_open_pores: bool = True
    real_signal_check: bool = True
    mean_window: int = 300
    mean_start_range: tuple = (-INF, INF)
    mean_end_range: tuple = (-INF, INF)
    max_obs_local_range: int = 5000
    local_range: tuple = (7.0, 35.0)
    adapter_mad_range: tuple = (3.0, 12.0)
    open_pore_min: float = 200.0        # find_open_pores default sig_range=(200.0, None)
    open_pore_min_obs_diff: int = 10    # find_open_pores default min_obs_diff
    # [mvs_polya]
    mvs_detect_check: bool = True
    pA_mean_window: int = 20
    pA_var_window: int = 100
    pA_var_range: tuple = (-INF, 30.0)
    median_shift_range: tuple = (5.0, INF)
    median_shift_window: int = 1000
    polyA_med_range: tuple = (-INF, INF)
    polyA_local_range: tuple = (-INF, INF)
    pA_mean_range: tuple = (-INF, INF)
    pA_mean_adapter_med_scale_range: tuple = (1.3, INF)
    # [med_shift]
    detect_med_shift: bool = False
    med_shift_window: int = 2000
    med_shift_range: tuple = (5.0, INF)

    @classmethod
    def from_spc(cls, spc):
        """From the reference's SigProcConfig object (golden generator only)."""
        def rng(t):
            lo, hi = t
            return (-INF if lo is None else float(lo), INF if hi is None else float(hi))

        if spc.mvs_polya.mvs_detect_overwrite:
            raise NotImplementedError("mvs_detect_overwrite = true is not restated")
        return cls(
            min_obs_adapter=int(spc.core.min_obs_adapter),
            detect_open_pores=bool(spc.real_range.detect_open_pores),
            real_signal_check=bool(spc.real_range.real_signal_check),
            mean_window=int(spc.real_range.mean_window),
            mean_start_range=rng(spc.real_range.mean_start_range),
            mean_end_range=rng(spc.real_range.mean_end_range),
            max_obs_local_range=int(spc.real_range.max_obs_local_range),
            local_range=rng(spc.real_range.local_range),
            adapter_mad_range=rng(spc.real_range.adapter_mad_range),
            mvs_detect_check=bool(spc.mvs_polya.mvs_detect_check),
            pA_mean_window=int(spc.mvs_polya.pA_mean_window),
            pA_var_window=int(spc.mvs_polya.pA_var_window),
            pA_var_range=rng(spc.mvs_polya.pA_var_range),
            median_shift_range=rng(spc.mvs_polya.median_shift_range),
            median_shift_window=int(spc.mvs_polya.median_shift_window),
            polyA_med_range=rng(spc.mvs_polya.polyA_med_range),
            polyA_local_range=rng(spc.mvs_polya.polyA_local_range),
            pA_mean_range=rng(spc.mvs_polya.pA_mean_range),
            pA_mean_adapter_med_scale_range=rng(spc.mvs_polya.pA_mean_adapter_med_scale_range),
            detect_med_shift=bool(spc.med_shift.detect_med_shift),
            med_shift_window=int(spc.med_shift.med_shift_window),
            med_shift_range=rng(spc.med_shift.med_shift_range),
        )


def in_range(val, lo, hi) -> bool:          # utils.py:16-26 (scalar branch)
    return bool(lo <= val <= hi)


def range_is_empty(r) -> bool:               # utils.py:29-36
    return r[0] == -INF and r[1] == INF


def move_mean(a: np.ndarray, window: int) -> np.ndarray:
    """bottleneck.move_mean as documented (see module docstring): float32 out, first window-1 NaN."""
    out = np.full(a.shape, np.nan, dtype=a.dtype)
    if 1 <= window <= a.size:
        out[window - 1:] = np.lib.stride_tricks.sliding_window_view(a.astype(np.float64), window).mean(axis=1)
    return out


def move_var(a: np.ndarray, window: int) -> np.ndarray:
    out = np.full(a.shape, np.nan, dtype=a.dtype)
    if 1 <= window <= a.size:
        out[window - 1:] = np.lib.stride_tricks.sliding_window_view(a.astype(np.float64), window).var(axis=1)
    return out


def find_open_pores(sig: np.ndarray, lo: float, min_obs_diff: int) -> np.ndarray:   # anomalies.py:16-35
    pos = np.flatnonzero((lo <= sig) & (sig <= np.inf))
    if pos.size > 1:
        valid = [int(pos[i]) for i in range(1, pos.size) if pos[i] - pos[i - 1] >= min_obs_diff]
        if not valid:
            valid = [int(pos[-1])]
        return np.array(valid, dtype=np.int64)
    return pos.astype(np.int64)


def _mvs_check(sig: np.ndarray, adapter_end: int, polya_end: int, cfg: ValidateConfig, mean_range):
    """mean_var_shift_polyA_check(..., return_values=True, less_signal_ok=False, windowed_stats=True), mvs.py:42-159.
    Returns (success, check_vector[5], mean, var, polya_med, polya_local_range, med_shift)."""
    failed = (False, np.zeros(5, dtype=bool), 0.0, 0.0, 0.0, 0.0, 0.0)
    n = sig.size
    if polya_end == 0 or adapter_end == 0 or polya_end < adapter_end or polya_end - adapter_end <= 2:
        return failed
    if n < adapter_end + cfg.median_shift_window:
        return failed
    seg = sig[adapter_end:polya_end]
    if polya_end - adapter_end <= cfg.pA_var_window + 2:
        polya_var = np.var(seg)
    else:
        polya_var = np.nanmedian(move_var(seg, cfg.pA_var_window))
    if polya_end - adapter_end <= cfg.pA_mean_window + 2:
        polya_mean = np.mean(seg)
    else:
        polya_mean = np.nanmedian(move_mean(seg, cfg.pA_mean_window))
    polya_med = np.median(seg)
    polya_local_range = np.subtract(*np.percentile(seg, (85, 15)))
    med_shift = np.median(sig[adapter_end:min(adapter_end + cfg.median_shift_window, n)]) - np.median(
        sig[max(adapter_end - cfg.median_shift_window, 0):adapter_end])
    vals = (float(polya_mean), float(polya_var), float(polya_med), float(polya_local_range), float(med_shift))
    cv = np.array([in_range(vals[0], *mean_range), in_range(vals[1], *cfg.pA_var_range),
                   in_range(vals[2], *cfg.polyA_med_range), in_range(vals[3], *cfg.polyA_local_range),
                   in_range(vals[4], *cfg.median_shift_range)])
    return (bool(cv.all()), cv) + vals


def validate_one(row: np.ndarray, full_signal_len: int, adapter_end: int, polya_topk, cfg: ValidateConfig,
                 adapter_start: int = 0, verdict_only: bool = False):
    """validate_boundaries(signal[:full_signal_len], Boundaries(adapter_start, adapter_end, topk[0], topk), spc,
    full_signal_len) as combined_detect_cnn calls it (combined.py:215-221), for one minibatch row.
    Returns dict(success, code, checks (bit i set = check i passed), adapter_start, adapter_end, polya_end,
    vals[N_VALS], n_open_pores)."""
    sig = np.asarray(row)[:full_signal_len]
    vals = np.full(N_VALS, np.nan)
    out = dict(success=False, code=OK, checks=0, adapter_start=int(adapter_start), adapter_end=int(adapter_end),
               polya_end=int(polya_topk[0]) if len(polya_topk) else 0, vals=vals, n_open_pores=0, parts=np.full(N_PART, np.nan))
    if np.isnan(sig).any():                                         # combined.py:416-418 (raises; caller records it)
        out["code"] = HAS_NAN
        return out
    a0, a1 = int(adapter_start), int(adapter_end)
    polya_best = out["polya_end"]
    code = OK
    adapter_med = adapter_mad = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if a1 == 0:                                                 # combined.py:452-458
            code = NO_ADAPTER
        else:
            adapter_med = float(np.median(sig[a0:a1]))
            adapter_mad = float(np.median(np.abs(sig[a0:a1] - adapter_med)))
            vals[V_ADAPTER_MED], vals[V_ADAPTER_MAD] = adapter_med, adapter_mad
        if code == OK and adapter_mad and not in_range(adapter_mad, *cfg.adapter_mad_range):    # :461-467
            code = ADAPTER_MAD
        if code == OK and cfg.detect_open_pores:                    # :469-481
            pores = find_open_pores(sig[a0:a1], cfg.open_pore_min, cfg.open_pore_min_obs_diff)
            out["n_open_pores"] = int(pores.size)
            out["open_pores"] = pores + a0                         # DetectResults.open_pores (None when the step is skipped)
            if pores.size > 0:
                a0 = int(pores[-1]) + a0
                if a1 - a0 < cfg.min_obs_adapter:
                    code = OPEN_PORE
        if code == OK and cfg.real_signal_check:                    # :483-497, real_range.py:34-63
            seg = sig[a0:a1]
            ok = False
            if len(seg) >= 2 * cfg.mean_window:
                mean_start = np.mean(seg[:cfg.mean_window])
                mean_end = np.mean(seg[-cfg.mean_window:])
                vals[V_REAL_MEAN_START], vals[V_REAL_MEAN_END] = float(mean_start), float(mean_end)
                if in_range(float(mean_start), *cfg.mean_start_range) and in_range(float(mean_end), *cfg.mean_end_range):
                    lr = np.subtract(*np.percentile(seg[-min(cfg.max_obs_local_range, len(seg)):], (85, 15)))
                    vals[V_REAL_LOCAL_RANGE] = float(lr)
                    ok = in_range(lr, *cfg.local_range)
            if not ok:
                code = REAL_RANGE
        if code == OK and cfg.mvs_detect_check:                     # :499-575
            if polya_best == 0:
                code = NO_POLYA
            else:
                if range_is_empty(cfg.pA_mean_range) and not range_is_empty(cfg.pA_mean_adapter_med_scale_range):
                    mr = np.array(cfg.pA_mean_adapter_med_scale_range) * adapter_med
                    mean_range = (mr[0], mr[1])
                elif range_is_empty(cfg.pA_mean_range):
                    raise ValueError("pA_mean_range is not specified")
                else:
                    mean_range = cfg.pA_mean_range
                for pe in polya_topk:
                    pe = int(pe)
                    if pe == 0:
                        break
                    res = _mvs_check(sig, a1, pe, cfg, mean_range)
                    vals[V_MVS_MEAN:V_MVS_MED_SHIFT + 1] = res[2:7]
                    if not res[0]:
                        # once a candidate failed `success` stays False (combined.py:540-541, 608-610): later
                        # candidates are still evaluated and overwrite fail_reason and the reported values
                        if res[2] == 0:
                            code, out["checks"] = MVS_NO_SIGNAL, 0
                        else:
                            code, out["checks"] = MVS_CHECKS, int(sum(1 << i for i in range(5) if res[1][i]))
                    if code == OK:
                        polya_best = pe
                        break
                    if verdict_only:      # NOT the reference's behaviour: report the first failing candidate and stop
                        break
        if code == OK and cfg.detect_med_shift:                     # :612-629
            shift = float(np.median(sig[a1:min(a1 + cfg.med_shift_window, full_signal_len)])
                          - np.median(sig[max(a1 - cfg.med_shift_window, 0):a1]))
            vals[V_ADAPTER_RNA_MED_SHIFT] = shift
            if not in_range(shift, *cfg.med_shift_range):
                code = MED_SHIFT
    out.update(success=code == OK, code=code, adapter_start=a0, adapter_end=a1, polya_end=polya_best)
    out["parts"] = partitions(row, full_signal_len, a0, a1, polya_best)
    return out


N_PART = 18   # adapter, polya, rna_preloaded x (start, len, mean, std, med, mad); NaN = None


def _partition_stats(sig: np.ndarray, start, end):
    """calc_partition_stats (signal_partitions.py:80-96)."""
    if end <= start:
        return [float(start)] + [np.nan] * 5
    seg = sig[start:end]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, std, med = float(np.mean(seg)), float(np.std(seg)), float(np.median(seg))
        mad = float(np.median(np.abs(seg - med)))
    return [float(start), float(end - start), mean, std, med, mad]


def partitions(row: np.ndarray, full_signal_len: int, adapter_start: int, adapter_end: int, polya_end: int) -> np.ndarray:
    """calc_partitions_from_vals(signal[:full_signal_len], adapter_start, adapter_end, polya_end) as validate_boundaries
    calls it (combined.py:631-636): float64 [N_PART]."""
    sig = np.asarray(row)[:full_signal_len]
    return np.array(_partition_stats(sig, adapter_start, adapter_end) + _partition_stats(sig, adapter_end, polya_end)
                    + _partition_stats(sig, polya_end, sig.size))


def fail_reason(code: int, checks: int):
    if code == MVS_CHECKS:
        return FAIL_REASON[code] + " ".join(n for i, n in enumerate(CHECK_NAMES) if not (checks >> i) & 1)
    return FAIL_REASON[code]


def validate_batch(signals: np.ndarray, full_lens, preds: np.ndarray, cfg: ValidateConfig, verdict_only: bool = False):
    """Batch form: signals [n, stride] float32 (NaN padded), full_lens [n], preds [n, 1+k] (cnn_detect's output:
    adapter end, poly(A) end candidates).  Returns arrays success u8[n], code i32[n], checks i32[n],
    bounds i64[n,3] (adapter_start, adapter_end, polya_end), vals f64[n,N_VALS], n_open_pores i32[n]."""
    n = signals.shape[0]
    success = np.zeros(n, np.uint8)
    code = np.zeros(n, np.int32)
    checks = np.zeros(n, np.int32)
    bounds = np.zeros((n, 3), np.int64)
    vals = np.full((n, N_VALS), np.nan)
    pores = np.zeros(n, np.int32)
    for i in range(n):
        r = validate_one(signals[i], int(full_lens[i]), int(preds[i, 0]), preds[i, 1:], cfg, verdict_only=verdict_only)
        success[i], code[i], checks[i] = r["success"], r["code"], r["checks"]
        bounds[i] = (r["adapter_start"], r["adapter_end"], r["polya_end"])
        vals[i] = r["vals"]
        pores[i] = r["n_open_pores"]
    return success, code, checks, bounds, vals, pores


def open_pores_batch(signals: np.ndarray, full_lens, preds: np.ndarray, cfg: ValidateConfig):
    """DetectResults.open_pores of every row (combined.py:469-477, 676): None or the int64 positions."""
    return [validate_one(signals[i], int(full_lens[i]), int(preds[i, 0]), preds[i, 1:], cfg).get("open_pores")
            for i in range(signals.shape[0])]


def synthetic_case(seed: int, stride: int = 12000, k: int = 5):
    """Deterministic synthetic minibatch rows + CNN-like boundary predictions that reach every branch of
    validate_boundaries (shared by the golden generator and the tests; only numpy's Generator is used)."""
    rng = np.random.default_rng(seed)
    row = np.full(stride, np.nan, dtype=np.float32)
    kind = seed % 18
    a_len = int(rng.integers(1500, 5000))
    p_len = int(rng.integers(30, 1500)) if kind != 5 else int(rng.integers(3, 110))
    r_len = int(rng.integers(800, 4000)) if kind != 6 else int(rng.integers(50, 900))
    total = min(stride, a_len + p_len + r_len)
    a_mean = rng.uniform(60, 95)
    a_sd = rng.uniform(4, 14) if kind != 7 else rng.uniform(0.5, 3.0)
    # adapter: piecewise-constant levels + noise
    lv = np.repeat(rng.normal(a_mean, a_sd, a_len // 20 + 1), 20)[:a_len]
    adapter = lv + rng.normal(0, 1.5, a_len)
    if kind == 16:            # bimodal adapter: MAD inside its range, 85-15 percentile range above local_range
        far = rng.random(a_len) < 0.42
        adapter = np.where(far, a_mean + rng.choice([-1.0, 1.0], a_len) * rng.uniform(25, 40, a_len), a_mean + rng.normal(0, 6.0, a_len))
    polya = rng.normal(a_mean * rng.uniform(1.15, 1.6), rng.uniform(1.0, 7.0), p_len)
    rl = np.repeat(rng.normal(a_mean * 1.1, 12.0, r_len // 12 + 1), 12)[:r_len]
    rna = rl + rng.normal(0, 2.0, r_len)
    sig = np.concatenate([adapter, polya, rna])[:total].astype(np.float32)
    if kind in (2, 3, 10):    # open-pore spikes inside the adapter
        npore = int(rng.integers(1, 4)) if kind != 10 else 1
        for _ in range(npore):
            p = int(rng.integers(0, a_len - (0 if kind == 3 else 1200)))
            w = int(rng.integers(1, 25))
            sig[p:p + w] = rng.uniform(200, 260, min(w, total - p)).astype(np.float32)
    if kind == 11:            # two isolated open-pore samples closer than min_obs_diff
        p = int(rng.integers(10, a_len - 1300))
        sig[p] = 230.0
        sig[p + 4] = 215.0
    if kind == 12:            # exactly one open-pore sample
        sig[int(rng.integers(10, a_len - 1300))] = 200.0
    if kind == 17:            # short read: the predicted adapter end lies beyond the signal
        total = int(rng.integers(300, 1500))
        sig = sig[:total]
    row[:total] = sig
    full_len = total if kind != 13 else total + int(rng.integers(1, 5000))   # read longer than the preloaded row
    if kind == 13:
        row[:] = np.resize(sig, stride)
    if kind == 14:            # NaN inside the valid part (exception path)
        row[int(rng.integers(0, total))] = np.nan
    a_end = a_len + int(rng.integers(-8, 9)) * 10 if kind != 1 else 0
    a_end = max(0, a_end)
    cands = [a_len + p_len + int(rng.integers(-5, 6)) * 10]
    for _ in range(k - 1):
        cands.append(a_len + int(rng.integers(1, max(2, (p_len + r_len) // 10))) * 10 if rng.random() < 0.6 else 0)
    nz = [c for c in cands if c > 0]
    cands = nz + [0] * (k - len(nz))
    if kind == 4:
        cands = [0] * k
    if kind == 8:             # candidate before the adapter end / within 2 samples of it
        cands[0] = a_end - 10 if rng.random() < 0.5 else a_end + 2
    if kind == 9:             # first candidate inside the RNA (fails), a later one at the true poly(A) end
        cands = [a_len + p_len + int(rng.integers(40, 80)) * 10, a_len + p_len] + cands[2:]
    preds = np.array([a_end] + cands[:k], dtype=np.int64)
    return row, np.int32(full_len), preds
