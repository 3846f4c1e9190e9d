"""Boundary validation and the CNN detection workflow — host-side mirror of the reference's
`adapted.detect.combined` (warpdemux/adapted/adapted/detect/combined.py) on top of the CUDA library
(include/wdx_b200.h: wdx_validate_*, wdx_cnn_*).

    validate_boundaries_batch(signals, full_signal_lens, preds, spc)   combined.py:409-683, whole minibatch
    combined_detect_cnn(batch_of_signals, full_signal_lens, model, spc) combined.py:198-306 -> List[DetectResults]

`spc` is the reference's `SigProcConfig` or anything with the same attributes (`ValidateConfig.from_spc`,
`LLRConfig.from_spc`).  Reads whose CNN boundaries fail validation are re-detected on the device like the reference
does on the CPU (combined.py:222-290: poly(A) re-detection on the CNN adapter end for short reads, then the full LLR
detector; `wdx_validate_set_llr`, csrc/llr_kernel.cuh) when the config's `cnn_boundaries.fallback_to_llr*` flags ask
for it.  All arithmetic is in warpdemux_b200/csrc/{validate,llr}_kernel.cuh; there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .. import _lib
from ..sharding import default_device
from . import cnn as _cnn

INF = float("inf")
N_VALS = 12
N_PART = 18
PORES_LD = 64     # WDX_VAL_PORES_LD
PART_FIELDS = tuple(f"{p}_{f}" for p in ("adapter", "polya", "rna_preloaded") for f in ("start", "len", "mean", "std", "med", "mad"))
FAIL_REASONS = {
    0: None,
    1: "No adapter detected (primary)",
    2: "adapter MAD check failed",
    3: "Open pore too close to boundary",
    4: "Real signal check failed",
    5: "No polya detected (primary)",
    6: "MVS polya check failed: not enough signal",
    7: "MVS polya check failed: ",
    8: "Median shift check failed",
    9: "Validate boundaries failed: Signal contains nan values",
    10: "MAD normalization failed: scale is 0",       # normalize.py:55-58, raised inside the fallback branch
    11: "LLR detection failed",
}
_CHECK_NAMES = ("mean", "var", "med", "range", "shift")
HAS_NAN = 9


def _rng(t):
    lo, hi = t
    return (-INF if lo is None else float(lo), INF if hi is None else float(hi))


@dataclass
class ValidateConfig:
    """The SigProcConfig fields validate_boundaries reads (defaults: adapted rna004_130bps@v0.2.4.toml)."""
    min_obs_adapter: int = 1000
    detect_open_pores: bool = True
    real_signal_check: bool = True
    mean_window: int = 300
    mean_start_range: tuple = (-INF, INF)
    mean_end_range: tuple = (-INF, INF)
    max_obs_local_range: int = 5000
    local_range: tuple = (7.0, 35.0)
    adapter_mad_range: tuple = (3.0, 12.0)
    open_pore_min: float = 200.0
    open_pore_min_obs_diff: int = 10
    mvs_detect_check: bool = True
    mvs_detect_overwrite: bool = False
    pA_mean_window: int = 20
    pA_var_window: int = 100
    pA_var_range: tuple = (-INF, 30.0)
    median_shift_range: tuple = (5.0, INF)
    median_shift_window: int = 1000
    polyA_med_range: tuple = (-INF, INF)
    polyA_local_range: tuple = (-INF, INF)
    pA_mean_range: tuple = (-INF, INF)
    pA_mean_adapter_med_scale_range: tuple = (1.3, INF)
    detect_med_shift: bool = False
    med_shift_window: int = 2000
    med_shift_range: tuple = (5.0, INF)
    primary_method: str = "cnn"

    @classmethod
    def from_spc(cls, spc) -> "ValidateConfig":
        if isinstance(spc, cls):
            return spc
        rr, mv, ms = spc.real_range, spc.mvs_polya, spc.med_shift
        return cls(
            min_obs_adapter=int(spc.core.min_obs_adapter),
            detect_open_pores=bool(rr.detect_open_pores), real_signal_check=bool(rr.real_signal_check),
            mean_window=int(rr.mean_window), mean_start_range=_rng(rr.mean_start_range), mean_end_range=_rng(rr.mean_end_range),
            max_obs_local_range=int(rr.max_obs_local_range), local_range=_rng(rr.local_range),
            adapter_mad_range=_rng(rr.adapter_mad_range),
            mvs_detect_check=bool(mv.mvs_detect_check), mvs_detect_overwrite=bool(mv.mvs_detect_overwrite),
            pA_mean_window=int(mv.pA_mean_window), pA_var_window=int(mv.pA_var_window), pA_var_range=_rng(mv.pA_var_range),
            median_shift_range=_rng(mv.median_shift_range), median_shift_window=int(mv.median_shift_window),
            polyA_med_range=_rng(mv.polyA_med_range), polyA_local_range=_rng(mv.polyA_local_range),
            pA_mean_range=_rng(mv.pA_mean_range), pA_mean_adapter_med_scale_range=_rng(mv.pA_mean_adapter_med_scale_range),
            detect_med_shift=bool(ms.detect_med_shift), med_shift_window=int(ms.med_shift_window),
            med_shift_range=_rng(ms.med_shift_range), primary_method=str(getattr(spc, "primary_method", "cnn")),
        )


@dataclass
class LLRConfig:
    """The SigProcConfig fields the LLR fallback reads (WarpDemuX rna004_130bps@v1.0.toml over adapted @v0.2.4)."""
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

    @classmethod
    def from_spc(cls, spc) -> Optional["LLRConfig"]:
        """None when the config has no llr_boundaries / cnn_boundaries sections to read the fallback from."""
        if isinstance(spc, cls):
            return spc
        core, cb, lb = getattr(spc, "core", None), getattr(spc, "cnn_boundaries", None), getattr(spc, "llr_boundaries", None)
        if core is None or cb is None or lb is None or not hasattr(cb, "fallback_to_llr"):
            return None
        return cls(max_obs_trace=int(core.max_obs_trace), min_obs_adapter=int(core.min_obs_adapter),
                   max_obs_adapter=int(core.max_obs_adapter), downscale_factor=int(core.downscale_factor),
                   sig_norm_outlier_thresh=float(core.sig_norm_outlier_thresh),
                   adapter_peak_prominence=float(lb.adapter_peak_prominence), adapter_peak_rel_height=float(lb.adapter_peak_rel_height),
                   adapter_peak_width=int(lb.adapter_peak_width), fallback_to_llr=bool(cb.fallback_to_llr),
                   fallback_to_llr_short_reads=bool(getattr(cb, "fallback_to_llr_short_reads", False)))


class _CLLRConfig(C.Structure):
    _fields_ = [("max_obs_trace", C.c_int32), ("min_obs_adapter", C.c_int32), ("max_obs_adapter", C.c_int32),
                ("downscale_factor", C.c_int32), ("sig_norm_outlier_thresh", C.c_double), ("adapter_peak_prominence", C.c_double),
                ("adapter_peak_rel_height", C.c_double), ("adapter_peak_width", C.c_int32), ("fallback_to_llr", C.c_int32),
                ("fallback_to_llr_short_reads", C.c_int32), ("reserved", C.c_int32)]


class _CValidateConfig(C.Structure):
    _fields_ = [
        ("min_obs_adapter", C.c_int32), ("detect_open_pores", C.c_int32), ("real_signal_check", C.c_int32),
        ("mean_window", C.c_int32), ("max_obs_local_range", C.c_int32), ("open_pore_min_obs_diff", C.c_int32),
        ("open_pore_min", C.c_double),
        ("mean_start_range", C.c_double * 2), ("mean_end_range", C.c_double * 2), ("local_range", C.c_double * 2),
        ("adapter_mad_range", C.c_double * 2),
        ("mvs_detect_check", C.c_int32), ("mvs_detect_overwrite", C.c_int32), ("pA_mean_window", C.c_int32),
        ("pA_var_window", C.c_int32), ("median_shift_window", C.c_int32), ("detect_med_shift", C.c_int32),
        ("med_shift_window", C.c_int32), ("reserved", C.c_int32),
        ("pA_var_range", C.c_double * 2), ("median_shift_range", C.c_double * 2), ("polyA_med_range", C.c_double * 2),
        ("polyA_local_range", C.c_double * 2), ("pA_mean_range", C.c_double * 2),
        ("pA_mean_adapter_med_scale_range", C.c_double * 2), ("med_shift_range", C.c_double * 2),
    ]


def _c_config(cfg: ValidateConfig) -> _CValidateConfig:
    c = _CValidateConfig()
    for name, _typ in _CValidateConfig._fields_:
        if name == "reserved":
            continue
        v = getattr(cfg, name)
        if isinstance(v, (tuple, list)):
            setattr(c, name, (C.c_double * 2)(float(v[0]), float(v[1])))
        elif isinstance(v, float):
            setattr(c, name, v)
        else:
            setattr(c, name, int(v))
    return c


@dataclass
class DetectResults:
    """adapted/container_types.py:14-77: the reference's fields in the reference's order (so `to_dict()` yields the
    columns of its detected_boundaries table), then this package's extras.  Not reproduced: `llr_trace` (dropped by the
    reference's writer) and the start-peak fields, which belong to a detector that is not on the GPU path."""
    success: bool

    signal_len: Optional[int] = None
    preloaded: Optional[int] = None

    adapter_start: Optional[int] = None
    adapter_end: Optional[int] = None
    adapter_len: Optional[int] = None
    adapter_mean: Optional[float] = None
    adapter_std: Optional[float] = None
    adapter_med: Optional[float] = None
    adapter_mad: Optional[float] = None

    polya_start: Optional[int] = None
    polya_end: Optional[int] = None
    polya_len: Optional[int] = None
    polya_mean: Optional[float] = None
    polya_std: Optional[float] = None
    polya_med: Optional[float] = None
    polya_mad: Optional[float] = None
    polya_candidates: Optional[np.ndarray] = None

    rna_preloaded_start: Optional[int] = None
    rna_preloaded_len: Optional[int] = None
    rna_preloaded_mean: Optional[float] = None
    rna_preloaded_std: Optional[float] = None
    rna_preloaded_med: Optional[float] = None
    rna_preloaded_mad: Optional[float] = None

    start_peak_idx: Optional[int] = None
    start_peak_pa: Optional[float] = None
    start_peak_next_max_idx: Optional[int] = None
    start_peak_next_max_pa: Optional[float] = None
    start_peak_open_pore_idx: Optional[int] = None

    adapter_rna_median_shift: Optional[float] = None

    llr_adapter_end: Optional[int] = None
    llr_polya_end: Optional[int] = None

    cnn_adapter_end: Optional[int] = None
    cnn_polya_end: Optional[int] = None

    start_peak_adapter_end: Optional[int] = None
    start_peak_polya_end: Optional[int] = None

    llr_trace: Optional[np.ndarray] = None

    mvs_adapter_end: Optional[int] = None
    mvs_detect_mean_at_loc: Optional[float] = None
    mvs_detect_var_at_loc: Optional[float] = None
    mvs_detect_polya_med: Optional[float] = None
    mvs_detect_polya_local_range: Optional[float] = None
    mvs_detect_med_shift: Optional[float] = None

    real_adapter_mean_start: Optional[float] = None
    real_adapter_mean_end: Optional[float] = None
    real_adapter_local_range: Optional[float] = None

    open_pores: Optional[np.ndarray] = None

    fail_reason: Optional[str] = None

    # extras of this package
    n_open_pores: int = 0
    needs_llr_fallback: bool = False     # validation failed and no device LLR fallback ran for this read
    detect_source: int = 0               # 0 = CNN boundaries, 1 = hail-mary poly(A) re-detection, 2 = full LLR detection

    def to_dict(self):
        d = dict(self.__dict__)
        for k in ("n_open_pores", "needs_llr_fallback", "detect_source"):
            d.pop(k, None)
        return d

    def update(self, d: dict):
        self.__dict__.update(d)


@dataclass
class ValidationBatch:
    success: np.ndarray        # uint8 [n]
    code: np.ndarray           # int32 [n] WDX_VAL_*
    checks: np.ndarray         # int32 [n] bit i set = MVS check i passed
    n_open_pores: np.ndarray   # int32 [n]
    bounds: np.ndarray         # int64 [n, 3] adapter_start, adapter_end, polya_end
    vals: np.ndarray           # float64 [n, N_VALS]
    kernel_ms: Optional[float] = field(default=None)
    parts: Optional[np.ndarray] = field(default=None)   # float64 [n, N_PART] partition statistics (PART_FIELDS), NaN = None
    open_pores: Optional[np.ndarray] = field(default=None)   # int32 [n, PORES_LD]: count (-1 = None), positions
    source: Optional[np.ndarray] = field(default=None)  # int32 [n] info[.][3]: bits 0-1 boundaries validated (0 cnn, 1 hail mary,
                                                        # 2 llr), bit 2 hail mary ran, bit 3 full LLR ran

    def fail_reason(self, i: int) -> Optional[str]:
        return fail_reason(int(self.code[i]), int(self.checks[i]))


def fail_reason(code: int, checks: int = 0) -> Optional[str]:
    if code == 7:
        return FAIL_REASONS[7] + " ".join(n for i, n in enumerate(_CHECK_NAMES) if not (checks >> i) & 1)
    return FAIL_REASONS[code]


class Validator:
    """Device handle of the validation step (`wdx_validate*`), created lazily in the process that uses it."""

    def __init__(self, spc=None, device: Optional[int] = None, verdict_only: bool = False, llr="auto"):
        """verdict_only: stop at the first failing poly(A) candidate — same `success` and boundaries, fail reason and
        mvs_* values of that candidate instead of the last one the reference goes on to evaluate.
        llr: an LLRConfig = re-detect the reads that fail on the device (combined.py:222-290); None = off; "auto" = what
        `spc` says (off for a bare ValidateConfig)."""
        self.cfg = ValidateConfig.from_spc(spc) if spc is not None else ValidateConfig()
        self.device = device
        self.verdict_only = bool(verdict_only)
        if isinstance(llr, str):
            llr = LLRConfig.from_spc(spc) if (spc is not None and not isinstance(spc, ValidateConfig)) else None
        if llr is not None and not (llr.fallback_to_llr or llr.fallback_to_llr_short_reads):
            llr = None
        self.llr = llr
        self._h = None

    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            c = _c_config(self.cfg)
            dev = default_device() if self.device is None else int(self.device)
            _lib.check(_lib.load().wdx_validate_create(C.byref(c), dev, C.byref(h)), "wdx_validate_create")
            _lib.check(_lib.load().wdx_validate_set_verdict_only(h, int(self.verdict_only)), "wdx_validate_set_verdict_only")
            if self.llr is not None:
                lc = _CLLRConfig()
                for name, _t in _CLLRConfig._fields_:
                    if name != "reserved":
                        v = getattr(self.llr, name)
                        setattr(lc, name, float(v) if isinstance(v, float) else int(v))
                _lib.check(_lib.load().wdx_validate_set_llr(h, C.byref(lc)), "wdx_validate_set_llr")
            self._h = h
        return self._h

    def run_raw(self, signals, n: int, stride: int, full_lens, preds, ld: int, success, info, bounds, vals=None, stream: int = 0,
                parts=None, open_pores=None):
        """Pointer-level call (numpy arrays, torch tensors or addresses; host or device memory)."""
        p = _cnn._ptr
        rc = _lib.load().wdx_validate_run_report(self._handle(), p(signals), int(n), int(stride), p(full_lens), p(preds), int(ld),
                                                 p(success), p(info), p(bounds), p(vals), p(parts), p(open_pores), stream or None)
        _lib.check(rc, "wdx_validate_run_report")

    def set_early(self, success_snapshot, event) -> None:
        """One-shot for the next run: `success` is copied to `success_snapshot` (device uint8 [n]) and `event` (a recorded
        torch.cuda.Event) is re-recorded right behind the first validation pass, before the LLR re-detection of the failed
        reads is queued (wdx_validate_set_early)."""
        ev = int(event.cuda_event) if event is not None else 0
        if event is not None and not ev:
            raise ValueError("record the event once before handing it over (torch creates the CUDA event lazily)")
        _lib.check(_lib.load().wdx_validate_set_early(self._handle(), _cnn._ptr(success_snapshot), ev or None), "wdx_validate_set_early")

    def enable_timing(self, on: bool = True):
        _lib.check(_lib.load().wdx_validate_enable_timing(self._handle(), int(on)), "wdx_validate_enable_timing")

    def last_kernel_ms(self) -> float:
        ms, k = C.c_double(), C.c_int()
        _lib.check(_lib.load().wdx_validate_last_kernel_ms(self._handle(), C.byref(ms), C.byref(k)), "wdx_validate_last_kernel_ms")
        return ms.value

    def validate(self, batch_of_signals, full_signal_lens, preds, partitions: bool = False) -> ValidationBatch:
        sig = _cnn._as_batch(batch_of_signals)
        n, stride = sig.shape
        lens = np.ascontiguousarray(np.minimum(np.asarray(full_signal_lens, dtype=np.int64), np.iinfo(np.int32).max), dtype=np.int32)
        pr = np.ascontiguousarray(preds, dtype=np.int64)
        if pr.ndim != 2 or pr.shape[0] != n or lens.shape != (n,):
            raise ValueError("preds must be [n, 1 + k] and full_signal_lens [n]")
        success = np.zeros(n, np.uint8)
        info = np.zeros((n, 4), np.int32)
        bounds = np.zeros((n, 3), np.int64)
        vals = np.full((n, N_VALS), np.nan)
        parts = np.full((n, N_PART), np.nan) if partitions else None
        pores = np.full((n, PORES_LD), -1, np.int32)
        if n:
            self.run_raw(sig, n, stride, lens, pr, pr.shape[1], success, info, bounds, vals, parts=parts, open_pores=pores)
        return ValidationBatch(success, info[:, 0].copy(), info[:, 1].copy(), info[:, 2].copy(), bounds, vals, parts=parts,
                               open_pores=pores, source=info[:, 3].copy())

    def close(self):
        if self._h is not None:
            _lib.load().wdx_validate_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self):
        return {"cfg": self.cfg, "device": self.device, "verdict_only": self.verdict_only, "llr": self.llr}

    def __setstate__(self, st):
        self.cfg, self.device, self._h = st["cfg"], st["device"], None
        self.verdict_only = st.get("verdict_only", False)
        self.llr = st.get("llr")


def validate_boundaries_batch(batch_of_signals, full_signal_lens, preds, spc, validator: Optional[Validator] = None,
                              partitions: bool = False) -> ValidationBatch:
    v = validator or Validator(spc)
    try:
        return v.validate(batch_of_signals, full_signal_lens, preds, partitions=partitions)
    finally:
        if validator is None:
            v.close()


def _opt(x: float) -> Optional[float]:
    return None if x != x else float(x)


def open_pores_array(row: Optional[np.ndarray]) -> Optional[np.ndarray]:
    """One row of `ValidationBatch.open_pores` -> what the reference stores in DetectResults.open_pores: None when
    the open-pore step did not run, else the int64 positions (combined.py:469-477)."""
    if row is None or row[0] < 0:
        return None
    m = int(row[0])
    if m > PORES_LD - 1:
        raise ValueError(f"{m} open pores in one adapter: more than the {PORES_LD - 1} the validation kernel lists")
    return row[1:1 + m].astype(np.int64)


def to_detect_results(vb: ValidationBatch, preds: np.ndarray, full_signal_lens, stride: int, primary_method: str = "cnn",
                      llr_ran: bool = False) -> List[DetectResults]:
    out = []
    for i in range(len(vb.success)):
        code = int(vb.code[i])
        if code in (HAS_NAN, 10, 11):   # the reference raises; combined_detect_cnn records str(e) (combined.py:293-294)
            out.append(DetectResults(success=False, fail_reason=FAIL_REASONS[code]))
            continue
        src = int(vb.source[i]) & 3 if vb.source is not None else 0
        fl = int(full_signal_lens[i])
        v = vb.vals[i]
        d = DetectResults(
            success=bool(vb.success[i]), signal_len=fl, preloaded=min(fl, int(stride)),
            adapter_start=int(vb.bounds[i, 0]), adapter_end=int(vb.bounds[i, 1]), polya_end=int(vb.bounds[i, 2]),
            polya_candidates=np.asarray(preds[i, 1:]).copy() if src == 0 else np.array([int(vb.bounds[i, 2])]),
            mvs_detect_mean_at_loc=_opt(v[5]), mvs_detect_var_at_loc=_opt(v[6]), mvs_detect_polya_med=_opt(v[7]),
            mvs_detect_polya_local_range=_opt(v[8]), mvs_detect_med_shift=_opt(v[9]), adapter_rna_median_shift=_opt(v[10]),
            # np.float32 like the reference's np.mean over float32 samples (real_range.py:47-48): a table whose rows all
            # carry the value gets a float32 column, which pandas rounds and prints differently from float64
            real_adapter_mean_start=None if v[2] != v[2] else np.float32(v[2]),
            real_adapter_mean_end=None if v[3] != v[3] else np.float32(v[3]), real_adapter_local_range=_opt(v[4]),
            n_open_pores=int(vb.n_open_pores[i]), fail_reason=vb.fail_reason(i),
            needs_llr_fallback=not bool(vb.success[i]) and not llr_ran, detect_source=src,
            open_pores=(open_pores_array(vb.open_pores[i]) if vb.open_pores is not None
                        else (np.array([int(vb.bounds[i, 0])]) if vb.n_open_pores[i] > 0 else None)),
        )
        if vb.parts is not None:
            for name, x in zip(PART_FIELDS, vb.parts[i]):
                if name == "adapter_start":
                    continue   # already set from the boundaries
                val = None if x != x else (int(x) if name.endswith(("_start", "_len")) else float(x))
                setattr(d, name, val)
        if src == 0:
            setattr(d, f"{primary_method}_adapter_end", int(preds[i, 0]))
            setattr(d, f"{primary_method}_polya_end", int(preds[i, 1]) if preds.shape[1] > 1 else 0)
        else:       # validated under spc_copy.primary_method = "llr" (combined.py:229-230, 638-641)
            d.llr_adapter_end, d.llr_polya_end = int(vb.bounds[i, 1]), int(vb.bounds[i, 2])
        out.append(d)
    return out


def combined_detect_cnn(batch_of_signals: np.ndarray, full_signal_lens: np.ndarray, model: "_cnn.BoundariesCNN", spc,
                        validator: Optional[Validator] = None, mode: Optional[str] = None, partitions: bool = True) -> List[DetectResults]:
    """CNN boundaries + validation + hail-mary / LLR re-detection of the reads that fail, for a minibatch, all on the GPU
    (combined.py:198-306).  The LLR branch runs when `spc` carries the reference's `cnn_boundaries.fallback_to_llr*` and
    `llr_boundaries` settings (or `validator` was built with an LLRConfig); otherwise failed reads come back with
    `needs_llr_fallback = True`."""
    sig = _cnn._as_batch(batch_of_signals)
    preds = _cnn.cnn_detect(sig, model, spc.cnn_boundaries, spc.core, mode=mode)
    v = validator or Validator(spc)
    try:
        vb = v.validate(sig, full_signal_lens, preds, partitions=partitions)
    finally:
        if validator is None:
            v.close()
    return to_detect_results(vb, preds, full_signal_lens, sig.shape[1], str(getattr(spc, "primary_method", "cnn")),
                             llr_ran=v.llr is not None)
