"""CPU oracle for the adapter / poly(A) boundary CNN — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the
product (warpdemux_b200/) never does.

Restates the reference's `cnn_detect` chain (the step just
before the fingerprint stage):

  prepare_data     warpdemux/adapted/adapted/detect/cnn.py:71-85
    downscale      adapted/detect/downscale.py:4-41   (zero-pad to a multiple of the factor, block mean, float32)
    nanmedian/MAD  cnn.py:81-84                        (float32, NaN -> SCORE_EXCL = -5)
  BoundariesCNN    cnn.py:16-52   Conv1d(1,64,7,s3,p3) ReLU Conv1d(64,64,7,p3) ReLU Conv1d(64,64,7,p3) ReLU
                                  ConvTranspose1d(64,2,7,s3,p3); plain PyTorch fp32 here
  cnn_predict      cnn.py:104-162 argmax of channel 0 over the adapter range; channel 1 masked before the
                                  adapter end, argmax, masked after it; scipy find_peaks(distance=5) on the
                                  FLATTENED batch; per read the k highest peaks (descending height, stable)
  cnn_detect       cnn.py:165-183 * downscale_factor + min_obs_adapter; values equal to min_obs_adapter -> 0

Parity status: pinned against the reference's own `prepare_data`, `cnn_score` and `cnn_detect`
imported unmodified and run in the build container on real reads of test_data/demux/4000_rna004.pod5
(oracle/make_golden_cnn.py -> tests/golden/cnn_detect_rna004.npz): prepared inputs bit-identical,
boundaries identical, scores within float32 summation-order noise of torch's CPU convolution.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

SCORE_EXCL = np.float32(-5.0)  # cnn.py:13


@dataclass
class CnnConfig:
    """core.* and cnn_boundaries.* values the chain reads (adapted/config/sig_proc.py:22-57)."""

    min_obs_adapter: int = 1000
    max_obs_adapter: int = 6500
    downscale_factor: int = 10
    polya_cand_k: int = 5
    peak_distance: int = 5  # literal in cnn.py:139


def load_weights_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def prepare_data(signals: np.ndarray, cfg: CnnConfig) -> np.ndarray:
    """float32[n, m] NaN-padded pA rows -> float32[n, T] normalised, downscaled CNN input."""
    sig = np.asarray(signals, dtype=np.float32)[:, cfg.min_obs_adapter:]
    n, width = sig.shape
    f = cfg.downscale_factor
    rem = width % f
    if rem:
        sig = np.concatenate([sig, np.zeros((n, f - rem), dtype=np.float32)], axis=1)
    ds = sig.reshape(n, -1, f).mean(axis=2)                       # float32 block means
    med = np.nanmedian(ds, axis=-1, keepdims=True)
    mad = np.nanmedian(np.abs(ds - med), axis=-1, keepdims=True)
    x = (ds - med) / mad
    # torch.Tensor.nan_to_num(nan=-5): NaN -> -5, +-inf -> +-float32 max
    fmax = np.finfo(np.float32).max
    x = np.where(np.isnan(x), SCORE_EXCL, x)
    x = np.where(np.isposinf(x), fmax, x)
    x = np.where(np.isneginf(x), -fmax, x)
    return x.astype(np.float32)


def cnn_scores(x: np.ndarray, w: Dict[str, np.ndarray], threads: int = 0) -> np.ndarray:
    """Plain PyTorch fp32 forward of BoundariesCNN: float32[n, T] -> float32[n, 2, T_out]."""
    import torch
    import torch.nn.functional as F

    if threads:
        torch.set_num_threads(threads)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    with torch.no_grad():
        h = t(x)[:, None, :]
        h = F.relu(F.conv1d(h, t(w["w0"]), t(w["b0"]), stride=3, padding=3))
        h = F.relu(F.conv1d(h, t(w["w1"]), t(w["b1"]), padding=3))
        h = F.relu(F.conv1d(h, t(w["w2"]), t(w["b2"]), padding=3))
        h = F.conv_transpose1d(h, t(w["w3"]), t(w["b3"]), stride=3, padding=3)
    return h.numpy()


def cnn_predict(scores: np.ndarray, cfg: CnnConfig) -> np.ndarray:
    """float32[n, 2, T] -> int64[n, 1 + k] downscaled positions (adapter end, k poly(A) end candidates)."""
    from scipy.signal import find_peaks

    s = np.array(scores, dtype=np.float32, copy=True)
    n, _, T = s.shape
    k = cfg.polya_cand_k
    span = (cfg.max_obs_adapter - cfg.min_obs_adapter) // cfg.downscale_factor
    a_end = np.argmax(s[:, 0, :span], axis=1)
    if k < 1:
        return np.stack([a_end, np.zeros(n, dtype=np.int64)], axis=1)
    pos = np.arange(T)
    ch1 = s[:, 1, :]
    ch1[pos[None, :] < a_end[:, None]] = SCORE_EXCL
    p_end = np.argmax(ch1, axis=1)
    if k == 1:
        return np.stack([a_end, p_end], axis=1)
    ch1[pos[None, :] > p_end[:, None]] = SCORE_EXCL
    flat = ch1.reshape(-1)
    peaks, _ = find_peaks(flat, distance=cfg.peak_distance)
    out = np.zeros((n, 1 + k), dtype=np.int64)
    out[:, 0] = a_end
    owner = peaks // T
    # The reference groups the sorted candidates by runs of equal read index and writes the i-th GROUP
    # to row i (cnn.py:147-158, np.split at the switches + enumerate): a read without any peak shifts
    # the groups of all later reads up by one row.  Restated as is - the GPU path must agree with it.
    row = 0
    for r in np.unique(owner):
        mine = peaks[owner == r]
        # descending height, equal heights in ascending position (np.lexsort is stable)
        order = np.argsort(-flat[mine].astype(np.float64), kind="stable")
        top = (mine[order] % T)[:k]
        out[row, 1:1 + len(top)] = top
        row += 1
    return out


def cnn_detect_from_scores(scores: np.ndarray, cfg: CnnConfig) -> np.ndarray:
    preds = (cnn_predict(scores, cfg) * cfg.downscale_factor + cfg.min_obs_adapter).astype(np.int64)
    preds[preds == cfg.min_obs_adapter] = 0
    return preds


def cnn_detect(signals: np.ndarray, w: Dict[str, np.ndarray], cfg: CnnConfig) -> np.ndarray:
    return cnn_detect_from_scores(cnn_scores(prepare_data(signals, cfg), w), cfg)
