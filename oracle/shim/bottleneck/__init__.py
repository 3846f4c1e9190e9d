"""Test-only stand-in for the `bottleneck` package (absent from this image), so that the
reference's adapter DETECTION code (warpdemux/adapted/adapted/detect/mvs.py:93-107,255-263,
367-372; adapter_start.py:21) can run when golden fixtures are generated from real reads.
Detection is upstream of the accelerated path: its boundaries are INPUTS of the fixtures.
Semantics of bottleneck.move_mean / move_var: trailing window, min_count = window, the first
window-1 outputs are NaN, ddof = 0; float32 in -> float32 out (float64 accumulation)."""
import numpy as np


def _windows(a, window):
    a = np.asarray(a)
    if a.ndim != 1:
        raise NotImplementedError("1-D only")
    if window < 1 or a.size < window:
        return None, a
    return np.lib.stride_tricks.sliding_window_view(a.astype(np.float64), window), a


def move_mean(a, window, min_count=None, axis=-1):
    w, a = _windows(a, window)
    out = np.full(a.shape, np.nan, dtype=a.dtype if a.dtype.kind == "f" else np.float64)
    if w is not None:
        out[window - 1:] = w.mean(axis=1)
    return out


def move_var(a, window, min_count=None, axis=-1, ddof=0):
    w, a = _windows(a, window)
    out = np.full(a.shape, np.nan, dtype=a.dtype if a.dtype.kind == "f" else np.float64)
    if w is not None:
        out[window - 1:] = w.var(axis=1, ddof=ddof)
    return out
