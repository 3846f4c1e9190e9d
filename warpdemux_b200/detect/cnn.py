"""Adapter / poly(A) boundary CNN — host-side mirror of the reference's
`adapted.detect.cnn` (warpdemux/adapted/adapted/detect/cnn.py) on top of the
CUDA library (include/wdx_b200.h: wdx_cnn_*).

Same names and argument meaning as the reference:

    load_cnn_model(path)                                   cnn.py:55-68
    cnn_detect(batch_of_signals, model, params, core)      cnn.py:165-183  -> int [n, 1 + k]
    cnn_detect_boundaries(...)                             cnn.py:186-203  -> List[Boundaries]
    cnn_score_batch(batch_of_signals, model, core)         prepare_data + cnn_score (cnn.py:71-101), raw scores

`params` / `core` are the reference's `CNNBoundariesConfig` / `CoreConfig`
objects or anything with the same attributes.  All arithmetic is in
warpdemux_b200/csrc/cnn_kernels.cuh and cnn_tc_kernel.cuh; there is no CPU
implementation in this package — without the CUDA library these calls raise.
Validation of the boundaries (adapted/detect/combined.py:409-683) and the LLR
fallback stay with the caller.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .. import _lib
from ..sharding import default_device

MODES = {"exact": 0, "fast": 1, "guarded": 2}
FLAG_NONFINITE, FLAG_RECOMPUTED, FLAG_CHAIN, FLAG_RANGE = 1, 2, 4, 8
_KEYS = ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3")
_SHAPES = {"w0": (64, 1, 7), "b0": (64,), "w1": (64, 64, 7), "b1": (64,), "w2": (64, 64, 7), "b2": (64,),
           "w3": (64, 2, 7), "b3": (2,)}
_TORCH_KEYS = {"0.weight": "w0", "0.bias": "b0", "2.weight": "w1", "2.bias": "b1", "4.weight": "w2", "4.bias": "b2",
               "6.weight": "w3", "6.bias": "b3"}


@dataclass
class CoreConfig:
    """adapted/config/sig_proc.py:22-30 (values of rna004_130bps@v0.2.4 as WarpDemuX overrides them)."""
    min_obs_adapter: int = 1000
    max_obs_adapter: int = 6500
    min_obs_polya: int = 100
    downscale_factor: int = 10
    max_obs_trace: int = 10000
    sig_norm_outlier_thresh: float = 5.0


@dataclass
class CNNBoundariesConfig:
    """adapted/config/sig_proc.py:33-57."""
    cnn_detect: bool = True
    model_name: str = "rna004_130bps@v0.2.4.pth"
    polya_cand_k: int = 5
    fallback_to_llr_short_reads: bool = True
    fallback_to_llr: bool = True


@dataclass
class Boundaries:
    """adapted/container_types.py:7-13."""
    adapter_start: int
    adapter_end: int
    polya_end: int
    polya_end_topk: Optional[np.ndarray] = None
    trace: Optional[np.ndarray] = None


class _CConfig(C.Structure):
    _fields_ = [("min_obs_adapter", C.c_int32), ("max_obs_adapter", C.c_int32), ("downscale_factor", C.c_int32),
                ("polya_cand_k", C.c_int32), ("channels", C.c_int32), ("kernel_size", C.c_int32)]


class BoundariesCNN:
    """Weights of the reference's `BoundariesCNN` (cnn.py:16-52) as host arrays (picklable); the device
    replica (`wdx_cnn*`) is created lazily per (core config, k) in the process that uses it."""

    def __init__(self, weights: Dict[str, np.ndarray], device: Optional[int] = None, mode: str = "guarded"):
        self.weights = {}
        for k in _KEYS:
            w = np.ascontiguousarray(weights[k], dtype=np.float32)
            if w.shape != _SHAPES[k]:
                raise ValueError(f"{k}: shape {w.shape}, expected {_SHAPES[k]} (BoundariesCNN(channels=64, kernel_size=7))")
            self.weights[k] = w
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.device, self.mode = device, mode
        self._handles: Dict[tuple, C.c_void_p] = {}

    def state_dict(self) -> Dict[str, np.ndarray]:
        inv = {v: k for k, v in _TORCH_KEYS.items()}
        return {inv[k]: v for k, v in self.weights.items()}

    def eval(self):
        return self

    def _handle(self, core, k: int):
        key = (int(core.min_obs_adapter), int(core.max_obs_adapter), int(core.downscale_factor), int(k))
        if key not in self._handles:
            lib = _lib.load()
            cfg = _CConfig(key[0], key[1], key[2], key[3], 64, 7)
            h = C.c_void_p()
            dev = default_device() if self.device is None else int(self.device)
            w = self.weights
            _lib.check(lib.wdx_cnn_create(C.byref(cfg), *[w[n].ctypes.data for n in _KEYS], dev, C.byref(h)), "wdx_cnn_create")
            self._handles[key] = h
        return self._handles[key]

    def close(self):
        for h in self._handles.values():
            _lib.load().wdx_cnn_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self):
        return {"weights": self.weights, "device": self.device, "mode": self.mode}

    def __setstate__(self, st):
        self.weights, self.device, self.mode, self._handles = st["weights"], st["device"], st["mode"], {}


def load_cnn_model(path: str, device: Optional[int] = None, mode: str = "guarded") -> BoundariesCNN:
    """Weights from an .npz (w0,b0..w3,b3) or from the reference's torch state dict (.pth, cnn.py:55-68)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Model weights not found at {path}")
    if path.endswith(".npz"):
        with np.load(path) as z:
            return BoundariesCNN({k: z[k] for k in _KEYS}, device=device, mode=mode)
    import torch  # plumbing only: reads the reference's checkpoint format

    sd = torch.load(path, weights_only=True, map_location="cpu")
    return BoundariesCNN({_TORCH_KEYS[k]: v.numpy() for k, v in sd.items()}, device=device, mode=mode)


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


def detect_raw(model: BoundariesCNN, core, k: int, signals, n: int, stride: int, preds, scores=None, flags=None,
               mode: Optional[str] = None, stream: int = 0) -> None:
    """Pointer-level call (numpy arrays, torch tensors or addresses; host or device memory)."""
    rc = _lib.load().wdx_cnn_detect(model._handle(core, k), _ptr(signals), int(n), int(stride), MODES[mode or model.mode],
                                    _ptr(preds), _ptr(scores), _ptr(flags), stream or None)
    _lib.check(rc, "wdx_cnn_detect")


def score_len(model: BoundariesCNN, core, k: int, stride: int):
    t_in, t_out = C.c_int32(), C.c_int32()
    _lib.check(_lib.load().wdx_cnn_score_len(model._handle(core, k), int(stride), C.byref(t_in), C.byref(t_out)), "wdx_cnn_score_len")
    return t_in.value, t_out.value


def enable_timing(model: BoundariesCNN, core, k: int, on: bool = True):
    _lib.check(_lib.load().wdx_cnn_enable_timing(model._handle(core, k), int(on)), "wdx_cnn_enable_timing")


def last_kernel_ms(model: BoundariesCNN, core, k: int):
    ms, n = C.c_double(), C.c_int()
    _lib.check(_lib.load().wdx_cnn_last_kernel_ms(model._handle(core, k), C.byref(ms), C.byref(n)), "wdx_cnn_last_kernel_ms")
    return ms.value, n.value


def _as_batch(batch_of_signals) -> np.ndarray:
    sig = np.asarray(batch_of_signals)
    if sig.ndim != 2:
        raise ValueError("batch_of_signals must be 2-D [n_reads, n_samples]")
    if sig.dtype != np.float32 or not sig.flags.c_contiguous:
        sig = np.ascontiguousarray(sig, dtype=np.float32)
    return sig


def cnn_detect(batch_of_signals: np.ndarray, model: BoundariesCNN, params, core_params, mode: Optional[str] = None,
               return_flags: bool = False):
    """int64 [n, 1 + polya_cand_k]: adapter end and poly(A) end candidates in samples, 0 = none (cnn.py:165-183)."""
    k = int(params.polya_cand_k)
    if k < 2:
        raise NotImplementedError("polya_cand_k < 2 is not on the GPU path (shipped configs use 5, 10, 15)")
    sig = _as_batch(batch_of_signals)
    n = sig.shape[0]
    preds = np.zeros((n, 1 + k), dtype=np.int64)
    flags = np.zeros(n, dtype=np.uint8)
    if n:
        detect_raw(model, core_params, k, sig, n, sig.shape[1], preds, flags=flags, mode=mode)
    return (preds, flags) if return_flags else preds


def cnn_score_batch(batch_of_signals: np.ndarray, model: BoundariesCNN, params, core_params, mode: Optional[str] = None):
    """Raw CNN output float32 [n, 2, To] for a minibatch (prepare_data + cnn_score, cnn.py:71-101), with the
    boundaries of the same call."""
    k = int(params.polya_cand_k)
    sig = _as_batch(batch_of_signals)
    n = sig.shape[0]
    _, t_out = score_len(model, core_params, k, sig.shape[1])
    preds = np.zeros((n, 1 + k), dtype=np.int64)
    scores = np.zeros((n, 2, t_out), dtype=np.float32)
    if n:
        detect_raw(model, core_params, k, sig, n, sig.shape[1], preds, scores=scores, mode=mode)
    return scores, preds


def prepare_data(batch_of_signals: np.ndarray, core_params, model: BoundariesCNN, k: int = 5) -> np.ndarray:
    """float32 [n, T]: downscaled, median/MAD-normalised CNN input, NaN -> -5 (cnn.py:71-85; the reference returns the
    same values as a torch tensor [n, 1, T])."""
    sig = _as_batch(batch_of_signals)
    n = sig.shape[0]
    t_in, _ = score_len(model, core_params, k, sig.shape[1])
    x = np.zeros((n, t_in), dtype=np.float32)
    if n:
        _lib.check(_lib.load().wdx_cnn_prepare(model._handle(core_params, k), sig.ctypes.data, n, sig.shape[1], x.ctypes.data, None),
                   "wdx_cnn_prepare")
    return x


def cnn_predict(scores: np.ndarray, model: BoundariesCNN, params, core_params, return_flags: bool = False):
    """int64 [n, 1 + k] downscaled positions from raw scores float32 [n, 2, T] (cnn.py:104-162)."""
    k = int(params.polya_cand_k)
    if k < 2:
        raise NotImplementedError("polya_cand_k < 2 is not on the GPU path")
    sc = np.ascontiguousarray(scores, dtype=np.float32)
    if sc.ndim != 3 or sc.shape[1] != 2:
        raise ValueError("scores must be [n, 2, T]")
    n = sc.shape[0]
    preds = np.zeros((n, 1 + k), dtype=np.int64)
    flags = np.zeros(n, dtype=np.uint8)
    if n:
        _lib.check(_lib.load().wdx_cnn_predict(model._handle(core_params, k), sc.ctypes.data, n, sc.shape[2], 0, preds.ctypes.data,
                                               flags.ctypes.data, None), "wdx_cnn_predict")
    return (preds, flags) if return_flags else preds


def cnn_detect_boundaries(batch_of_signals: np.ndarray, model: BoundariesCNN, params, core_params,
                          mode: Optional[str] = None) -> List[Boundaries]:
    """cnn.py:186-203."""
    preds = cnn_detect(batch_of_signals, model, params, core_params, mode=mode)
    return [Boundaries(adapter_start=0, adapter_end=int(p[0]), polya_end=int(p[1]), polya_end_topk=p[1:].copy()) for p in preds]
