"""Fingerprint extraction — host-side mirror of the reference's
`warpdemux/sig_proc.py`: `detect_results_to_fpt` (sig_proc.py:394-605), both the
plain branch every shipped DTW-SVM model uses and the consensus-guided barcode
refinement of the tRNA configurations (sig_proc.py:257-378, 451-521).

The reference calls `detect_results_to_fpt` once per read inside a Python loop
(file_proc.py:418-428).  Here a whole minibatch goes to the GPU in one call of
`wdx_fp_extract` (include/wdx_b200.h): `batch_detect_results_to_fpt` is the
batched form, `detect_results_to_fpt` keeps the reference's per-read signature
on top of it.  All arithmetic is in warpdemux_b200/csrc/fingerprint_kernel.cuh;
there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .sharding import default_device

FP_OK, FP_FAIL_SEGMENTATION, FP_FAIL_DETECT, FP_FAIL_NORMALIZE, FP_FAIL_TOO_LONG, FP_FAIL_CONSENSUS = 0, 1, 2, 3, 4, 5
FP_NONFINITE = 6   # host-side status: fingerprint extracted, but a decision value was NaN / inf (sklearn raises there)
# the ONE table of fail reasons of the fingerprint stage, in the reference's words (file_proc / detect mirrors use it too)
FAIL_REASON = {
    FP_FAIL_CONSENSUS: "consensus query outlier",               # sig_proc.py:500-521
    FP_FAIL_SEGMENTATION: "event segmentation failed",          # sig_proc.py:472-479, 537-544
    FP_FAIL_DETECT: "detection failed",                         # sig_proc.py:400-407 reports detect_results.fail_reason instead
    FP_FAIL_NORMALIZE: "segment normalization failed: Signal contains NaN values.",   # sig_proc.py:553-560 + :100-103
    FP_FAIL_TOO_LONG: "adapter slice exceeds the GPU shared-memory limit",            # no counterpart in the reference
    FP_NONFINITE: "Input contains NaN.",                        # sklearn's check_array inside SVC.predict_proba
}
_FAIL_REASON = FAIL_REASON


# --------------------------------------------------------------------------
# containers (field-compatible with the reference's dataclasses)
# --------------------------------------------------------------------------
@dataclass
class DetectResults:
    """The fields of `adapted.container_types.DetectResults` this path reads
    (adapted/container_types.py:30-84); any object with these attributes works."""
    success: bool = True
    adapter_start: Optional[int] = None
    adapter_end: Optional[int] = None
    fail_reason: Optional[str] = None

    def to_dict(self):
        return {**self.__dict__}


@dataclass
class ReadResult:
    """`warpdemux.sig_proc.ReadResult` (sig_proc.py:26-62 on top of
    adapted/container_types.py:90-105)."""
    read_id: Optional[str] = None
    success: bool = True
    fail_reason: Optional[str] = None
    detect_results: Optional[Any] = None
    barcode_fpt: Optional[np.ndarray] = None
    dwell_times: Optional[np.ndarray] = None
    adapter_dt_med: Optional[float] = None
    adapter_dt_mad: Optional[float] = None
    adapter_event_mean: Optional[float] = None
    adapter_event_std: Optional[float] = None
    adapter_event_med: Optional[float] = None
    adapter_event_mad: Optional[float] = None
    seg_cons_query_start: Optional[int] = None
    seg_cons_query_end: Optional[int] = None
    sig_barcode_start: Optional[int] = None

    def to_summary_dict(self) -> Dict[str, Any]:
        d = self.detect_results.to_dict() if self.detect_results is not None and hasattr(self.detect_results, "to_dict") else {}
        d.pop("fail_reason", None)
        out = {"read_id": self.read_id, **d, "fail_reason": self.fail_reason}
        for key in ("adapter_dt_med", "adapter_dt_mad", "adapter_event_mean", "adapter_event_std", "adapter_event_med",
                    "adapter_event_mad", "seg_cons_query_start", "seg_cons_query_end", "sig_barcode_start"):
            out[key] = getattr(self, key)
        return out

    def set_read_id(self, read_id: str):
        self.read_id = read_id


@dataclass(frozen=True)
class FingerprintConfig:
    """The SigProcConfig keys the path reads; defaults = rna004_130bps@v1.0
    (config/config_files/rna004_130bps@v1.0.toml:5-14,
    adapted/config/config_files/rna004_130bps@v0.2.4.toml:7)."""
    padding: int = 100
    outlier_thresh: float = 5.0
    min_obs_per_base: int = 6
    running_stat_width: int = 12
    num_events: int = 110
    barcode_num_events: int = 25      # events kept (barcode_num_events, or barcode_num_events[1] with a consensus)
    max_slice_len: int = 0
    long_slice_len: int = 0           # > max_slice_len: second pass for the rare longer slices (wdx_fp_set_long_slice_len)
    numpy1_promotion: bool = False    # winsorisation bounds as numpy < 2 forms them (wdx_fp_set_numpy1_promotion)
    # consensus-guided barcode refinement (segmentation.consensus_refinement; rna004_130bps@v1.0_tRNA.toml:13-29)
    consensus: Optional[Tuple[float, ...]] = None   # warpdemux._consensus.ALL[consensus_model]; None = off
    barcode_segm_events: int = 25                    # barcode_num_events[0]
    consensus_penalty: float = 1.5
    consensus_psi: Tuple[int, int, int, int] = (5, 0, 40, 0)
    consensus_ub_start: int = 18
    consensus_lb_end: int = 69
    consensus_ub_end: int = 97

    @classmethod
    def trna(cls, consensus, **kw) -> "FingerprintConfig":
        """rna004_130bps@v1.0_tRNA.toml (WDX4_tRNA / WDX4b_tRNA) with the given consensus query."""
        base = dict(min_obs_per_base=9, running_stat_width=18, num_events=120, barcode_num_events=25)
        base.update(kw)
        return cls(consensus=tuple(float(v) for v in np.asarray(consensus).reshape(-1)), **base)

    @classmethod
    def from_spc(cls, spc, consensus_query=None) -> "FingerprintConfig":
        """From a reference `SigProcConfig` (or any object with the same attribute
        tree).  Raises for settings the GPU path does not implement."""
        seg, ext = spc.segmentation, spc.sig_extract
        if getattr(seg, "consensus_refinement", False):
            if getattr(seg, "refinement_optimal_cpts", False):
                raise NotImplementedError("refinement_optimal_cpts (ruptures KernelCPD) is not on the GPU path")
            if getattr(seg, "consensus_subseq_match_normalization", "mean") != "mean":
                raise NotImplementedError("consensus_subseq_match_normalization must be 'mean'")
            if consensus_query is None or not np.size(consensus_query):
                raise ValueError("consensus_refinement needs the consensus query (warpdemux._consensus.ALL[consensus_model])")
            nb = seg.barcode_num_events
            if isinstance(nb, (int, np.integer)):        # sig_proc.py:453-458
                raise ValueError("barcode_num_events is an integer in consensus refinement mode, use a tuple instead")
            psi = tuple(int(v) for v in seg.consensus_subseq_match_psi)
            if len(psi) != 4 or psi[1] or psi[3]:
                raise NotImplementedError("consensus_subseq_match_psi must be (begin_query, 0, begin_series, 0)")
            if getattr(ext, "normalization", "none") != "none" or getattr(seg, "normalization", "mean") != "mean" \
                    or getattr(seg, "accept_less_cpts", False):
                raise NotImplementedError("only normalization none/mean and accept_less_cpts=false are on the GPU path")
            return cls(padding=int(ext.padding), outlier_thresh=float(spc.core.sig_norm_outlier_thresh),
                       min_obs_per_base=int(seg.min_obs_per_base), running_stat_width=int(seg.running_stat_width),
                       num_events=int(seg.num_events), barcode_num_events=int(nb[1]),
                       consensus=tuple(float(v) for v in np.asarray(consensus_query).reshape(-1)),
                       barcode_segm_events=int(nb[0]), consensus_penalty=float(seg.consensus_subseq_match_penalty),
                       consensus_psi=psi, consensus_ub_start=int(seg.consensus_subseq_match_ub_start),
                       consensus_lb_end=int(seg.consensus_subseq_match_lb_end),
                       consensus_ub_end=int(seg.consensus_subseq_match_ub_end))
        if getattr(ext, "normalization", "none") != "none":
            raise NotImplementedError("sig_extract.normalization must be 'none'")
        if getattr(seg, "normalization", "mean") != "mean":
            raise NotImplementedError("segmentation.normalization must be 'mean'")
        if getattr(seg, "accept_less_cpts", False):
            raise NotImplementedError("accept_less_cpts=True is not on the GPU path")
        nb = seg.barcode_num_events
        if not isinstance(nb, (int, np.integer)):
            raise NotImplementedError("barcode_num_events must be an int without consensus refinement")
        return cls(padding=int(ext.padding), outlier_thresh=float(spc.core.sig_norm_outlier_thresh),
                   min_obs_per_base=int(seg.min_obs_per_base), running_stat_width=int(seg.running_stat_width),
                   num_events=int(seg.num_events), barcode_num_events=int(nb))


class _CConfig(C.Structure):
    _fields_ = [("padding", C.c_int32), ("outlier_thresh", C.c_double), ("min_obs_per_base", C.c_int32),
                ("running_stat_width", C.c_int32), ("num_events", C.c_int32), ("barcode_num_events", C.c_int32),
                ("max_slice_len", C.c_int32)]


class _CConsensus(C.Structure):
    _fields_ = [("query", C.c_void_p), ("query_len", C.c_int32), ("barcode_segm_events", C.c_int32),
                ("penalty", C.c_double), ("psi_query_begin", C.c_int32), ("psi_series_begin", C.c_int32),
                ("ub_start", C.c_int32), ("lb_end", C.c_int32), ("ub_end", C.c_int32)]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


@dataclass
class FingerprintBatch:
    fpt: np.ndarray                 # float64 [n, barcode_num_events], NaN rows where status != 0
    status: np.ndarray              # int32 [n]
    dwell: Optional[np.ndarray] = None   # int64 [n, barcode_num_events]
    stats: Optional[np.ndarray] = None   # float64 [n, 6]
    cons: Optional[np.ndarray] = None    # int32 [n, 3] seg_cons_query_start, seg_cons_query_end, sig_barcode_start


class Fingerprinter:
    """Owner of a `wdx_fp*` handle on one GPU (created lazily, so the object can
    be built in a parent process and used in workers)."""

    def __init__(self, config: FingerprintConfig = FingerprintConfig(), device: Optional[int] = None):
        self.config = config
        self.device = device
        self._h = None

    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            c = self.config
            cc = _CConfig(c.padding, c.outlier_thresh, c.min_obs_per_base, c.running_stat_width, c.num_events,
                          c.barcode_num_events, c.max_slice_len)
            h = C.c_void_p()
            dev = default_device() if self.device is None else int(self.device)
            _lib.check(lib.wdx_fp_create(C.byref(cc), dev, C.byref(h)), "wdx_fp_create")
            if c.numpy1_promotion:
                _lib.check(lib.wdx_fp_set_numpy1_promotion(h, 1), "wdx_fp_set_numpy1_promotion")
            if c.long_slice_len:
                rc = lib.wdx_fp_set_long_slice_len(h, int(c.long_slice_len))
                if rc != 0:
                    lib.wdx_fp_destroy(h)
                    _lib.check(rc, "wdx_fp_set_long_slice_len")
            if c.consensus is not None:
                q = np.ascontiguousarray(c.consensus, dtype=np.float64)
                cs = _CConsensus(q.ctypes.data, q.size, c.barcode_segm_events, c.consensus_penalty, c.consensus_psi[0],
                                 c.consensus_psi[2], c.consensus_ub_start, c.consensus_lb_end, c.consensus_ub_end)
                rc = lib.wdx_fp_set_consensus(h, C.byref(cs))
                if rc != 0:
                    lib.wdx_fp_destroy(h)
                    _lib.check(rc, "wdx_fp_set_consensus")
            self._h = h
        return self._h

    def close(self):
        if self._h is not None:
            _lib.load().wdx_fp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self):
        return {"config": self.config, "device": self.device}

    def __setstate__(self, st):
        self.config, self.device, self._h = st["config"], st["device"], None

    def enable_timing(self, on: bool = True):
        _lib.check(_lib.load().wdx_fp_enable_timing(self._handle(), int(on)), "wdx_fp_enable_timing")

    def last_kernel_ms(self):
        ms, n = C.c_double(), C.c_int()
        _lib.check(_lib.load().wdx_fp_last_kernel_ms(self._handle(), C.byref(ms), C.byref(n)), "wdx_fp_last_kernel_ms")
        return ms.value, n.value

    # -- pointer-level calls (numpy arrays, torch tensors or addresses) --------
    def set_resume_status(self, status: int) -> None:
        """One-shot: the fingerprint pass of the next extract_raw / predict_raw only revisits the reads whose `status` entry
        (device array written by an earlier call over the same batch) equals `status` (wdx_fp_set_resume_status)."""
        _lib.check(_lib.load().wdx_fp_set_resume_status(self._handle(), int(status)), "wdx_fp_set_resume_status")

    def extract_raw(self, signals, n, stride, adapter_start, adapter_end, fpt, status, sig_len=None, detect_ok=None,
                    dwell=None, stats=None, clip_in_place=False, stream: int = 0, cons=None) -> None:
        rc = _lib.load().wdx_fp_extract_ex(self._handle(), _ptr(signals), int(n), int(stride), _ptr(sig_len),
                                           _ptr(adapter_start), _ptr(adapter_end), _ptr(detect_ok), int(clip_in_place),
                                           _ptr(fpt), _ptr(dwell), _ptr(stats), _ptr(status), _ptr(cons), stream or None)
        _lib.check(rc, "wdx_fp_extract_ex")

    def predict_raw(self, device_model, signals, n, stride, adapter_start, adapter_end, mode, labels, status,
                    conf=None, prob=None, flags=None, fpt=None, sig_len=None, detect_ok=None, stream: int = 0) -> None:
        rc = _lib.load().wdx_fp_predict(self._handle(), device_model._h, _ptr(signals), int(n), int(stride),
                                        _ptr(sig_len), _ptr(adapter_start), _ptr(adapter_end), _ptr(detect_ok),
                                        int(mode), _ptr(labels), _ptr(conf), _ptr(prob), _ptr(flags), _ptr(fpt),
                                        _ptr(status), stream or None)
        _lib.check(rc, "wdx_fp_predict")

    # -- numpy in, numpy out -----------------------------------------------------
    @staticmethod
    def _prep(signals, adapter_start, adapter_end, sig_len, detect_ok):
        signals = np.asarray(signals)
        if signals.ndim == 1:
            signals = signals.reshape(1, -1)
        if signals.dtype != np.float32 or not signals.flags.c_contiguous:
            signals = np.ascontiguousarray(signals, dtype=np.float32)
        n = signals.shape[0]
        a0 = np.ascontiguousarray(np.asarray(adapter_start).reshape(-1), dtype=np.int64)
        a1 = np.ascontiguousarray(np.asarray(adapter_end).reshape(-1), dtype=np.int64)
        if a0.size != n or a1.size != n:
            raise ValueError("adapter_start / adapter_end must have one entry per signal row")
        sl = None if sig_len is None else np.ascontiguousarray(np.asarray(sig_len).reshape(-1), dtype=np.int32)
        ok = None if detect_ok is None else np.ascontiguousarray(np.asarray(detect_ok).reshape(-1), dtype=np.uint8)
        return signals, n, a0, a1, sl, ok

    def extract(self, signals, adapter_start, adapter_end, sig_len=None, detect_ok=None, want_dwell: bool = True,
                want_stats: bool = True, clip_in_place: bool = False) -> FingerprintBatch:
        """signals float32 [n, m] (rows may be NaN-padded at the end)."""
        given = signals
        signals, n, a0, a1, sl, ok = self._prep(signals, adapter_start, adapter_end, sig_len, detect_ok)
        if clip_in_place and signals is not given:
            raise ValueError("clip_in_place needs a C-contiguous float32 array")
        nb = self.config.barcode_num_events
        fpt = np.full((n, nb), np.nan, dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        dwell = np.zeros((n, nb), dtype=np.int64) if want_dwell else None
        stats = np.full((n, 6), np.nan, dtype=np.float64) if want_stats else None
        cons = np.zeros((n, 3), dtype=np.int32) if self.config.consensus is not None else None
        if n:
            self.extract_raw(signals, n, signals.shape[1], a0, a1, fpt, status, sig_len=sl, detect_ok=ok, dwell=dwell,
                             stats=stats, clip_in_place=clip_in_place, cons=cons)
        return FingerprintBatch(fpt=fpt, status=status, dwell=dwell, stats=stats, cons=cons)

    def extract_and_predict(self, model, signals, adapter_start, adapter_end, sig_len=None, detect_ok=None,
                            mode: Optional[str] = None, want_fpt: bool = False):
        """The fused minibatch step (file_proc.py:418-450): signals -> barcode calls;
        the fingerprints stay on the GPU.  `model` is a warpdemux_b200 `DTW_SVM`.
        Returns (labels int64[n], prob float64[n,k], conf float64[n], status int32[n][, fpt])."""
        signals, n, a0, a1, sl, ok = self._prep(signals, adapter_start, adapter_end, sig_len, detect_ok)
        dm = model._device_model()
        k = model.params.k
        labels = np.full(n, -1, dtype=np.int64)
        conf = np.full(n, np.nan)
        prob = np.full((n, k), np.nan)
        flags = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        fpt = np.full((n, self.config.barcode_num_events), np.nan)    # also the source of the overflow repair below
        if n:
            self.predict_raw(dm, signals, n, signals.shape[1], a0, a1, _lib.MODES[mode or model.mode], labels, status,
                             conf=conf, prob=prob, flags=flags, fpt=fpt, sig_len=sl, detect_ok=ok)
            over = np.flatnonzero((flags & _lib.FLAG_GUARD_OVERFLOW) != 0)
            if over.size:      # more guard-band reads than the re-run list of a launch holds: redo them in EXACT_F64
                l2, p2, c2, _ = dm.predict(fpt[over], mode="exact")
                labels[over], prob[over], conf[over] = l2, p2, c2
            # non-finite decision values with a valid fingerprint: the reference's predict raises; report the read failed
            status[(status == 0) & ((flags & _lib.FLAG_NONFINITE) != 0)] = FP_NONFINITE
        out = (labels, prob, conf, status)
        return out + (fpt,) if want_fpt else out


# --------------------------------------------------------------------------
# reference-shaped functions
# --------------------------------------------------------------------------
_default_fp: Dict[Any, Fingerprinter] = {}


def _fingerprinter_for(spc, consensus_query=None) -> Fingerprinter:
    cfg = spc if isinstance(spc, FingerprintConfig) else FingerprintConfig.from_spc(spc, consensus_query)
    key = (cfg, default_device())
    if key not in _default_fp:
        _default_fp[key] = Fingerprinter(cfg)
    return _default_fp[key]


def _read_result(b: FingerprintBatch, r: int, dr) -> ReadResult:
    st = int(b.status[r])
    if st == FP_FAIL_DETECT:  # sig_proc.py:400-407
        return ReadResult(success=False, fail_reason=getattr(dr, "fail_reason", None), barcode_fpt=np.array([]),
                          dwell_times=np.array([]), detect_results=dr)
    if st not in (FP_OK, FP_FAIL_CONSENSUS):
        return ReadResult(success=False, fail_reason=_FAIL_REASON.get(st, "unknown"), barcode_fpt=np.array([]),
                          dwell_times=np.array([]), detect_results=dr)
    s = b.stats[r]
    extra = dict(adapter_dt_med=float(s[0]), adapter_dt_mad=float(s[1]), adapter_event_mean=float(s[2]),
                 adapter_event_std=float(s[3]), adapter_event_med=float(s[4]), adapter_event_mad=float(s[5]))
    if b.cons is not None:
        extra.update(seg_cons_query_start=int(b.cons[r, 0]), seg_cons_query_end=int(b.cons[r, 1]),
                     sig_barcode_start=int(b.cons[r, 2]))
    if st == FP_FAIL_CONSENSUS:  # sig_proc.py:500-521: reported with the adapter statistics
        return ReadResult(success=False, fail_reason=_FAIL_REASON[st], barcode_fpt=np.array([]),
                          dwell_times=np.array([]), detect_results=dr, **extra)
    return ReadResult(success=True, fail_reason="", barcode_fpt=b.fpt[r].copy(), dwell_times=b.dwell[r].copy(),
                      detect_results=dr, **extra)


def batch_detect_results_to_fpt(signals: np.ndarray, spc, detect_results: Sequence[Any],
                                sig_len: Optional[Sequence[int]] = None,
                                clip_in_place: bool = False, consensus_query=None) -> List[ReadResult]:
    """Batched `detect_results_to_fpt`: signals float32 [n, m] (NaN-padded rows
    as `file_proc.yield_signals_from_pod5` builds them, file_proc.py:227-279),
    one DetectResults per row -> one ReadResult per row.  Like the reference's
    worker (file_proc.py:418-428) each padded row IS the signal; pass `sig_len`
    (samples per row) only to get the "signal without NaNs" behaviour instead."""
    fp = _fingerprinter_for(spc, consensus_query)
    n = len(detect_results)
    ok = np.array([bool(d.success) for d in detect_results], dtype=np.uint8)
    a0 = np.array([d.adapter_start if (d.success and d.adapter_start is not None) else 0 for d in detect_results],
                  dtype=np.int64)
    a1 = np.array([d.adapter_end if (d.success and d.adapter_end is not None) else 0 for d in detect_results],
                  dtype=np.int64)
    signals = np.asarray(signals)
    if signals.ndim == 1:
        signals = signals.reshape(1, -1)
    if signals.shape[0] != n:
        raise ValueError("one DetectResults per signal row is required")
    b = fp.extract(signals, a0, a1, sig_len=sig_len, detect_ok=ok, clip_in_place=clip_in_place)
    return [_read_result(b, r, detect_results[r]) for r in range(n)]


def detect_results_to_fpt(calibrated_signal: np.ndarray, spc, detect_results,
                          consensus_query: np.ndarray = np.array([])) -> ReadResult:
    """Per-read signature of the reference (sig_proc.py:394-399).  The signal must
    not contain NaNs (file_proc.py:194).  Like the reference, the winsorised
    adapter slice is written back into `calibrated_signal` when it is a
    C-contiguous float32 array."""
    sig = np.asarray(calibrated_signal)
    in_place = sig.dtype == np.float32 and sig.ndim == 1 and sig.flags.c_contiguous and sig.flags.writeable
    return batch_detect_results_to_fpt(sig.reshape(1, -1), spc, [detect_results], clip_in_place=in_place,
                                       consensus_query=consensus_query if np.size(consensus_query) else None)[0]
