"""Host-side helpers with the reference's names (warpdemux/models/utils.py).

`predictions_to_df` formats what the device already decided; `process_probs` /
`confidence_margin` are kept for API parity (e.g. re-thresholding stored
probabilities) — the predict path itself takes labels and margins from the
CUDA finishing kernel, not from here.
"""
from typing import Optional, Tuple

import numpy as np
import pandas as pd


def confidence_margin(npa: np.ndarray) -> np.ndarray:
    """top1 - top2 of each row (reference: models/utils.py:19-22)."""
    part = np.partition(npa, npa.shape[1] - 2, axis=1)
    return part[:, -1] - part[:, -2]


def predictions_to_df(y_pred: np.ndarray, y_prob: np.ndarray, conf: np.ndarray, label_mapper: dict) -> pd.DataFrame:
    """Same columns, order and rounding as the reference (models/utils.py:36-43):
    predicted_barcode, confidence_score (3 dp), p{label:02d} per class (4 dp)."""
    cols = {"predicted_barcode": y_pred, "confidence_score": conf.round(3)}
    for i in range(y_prob.shape[1]):
        cols[f"p{label_mapper[i]:02d}"] = y_prob[:, i].round(4)
    return pd.DataFrame(cols)


def process_probs(y_prob: np.ndarray, label_mapper: dict,
                  thresholds: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Probabilities -> (labels, confidence) (reference: models/utils.py:45-61)."""
    pred_idx = np.argmax(y_prob, axis=1)
    lut = np.array([label_mapper[i] for i in range(y_prob.shape[1])])
    pred = lut[pred_idx]
    conf = confidence_margin(y_prob)
    if thresholds is not None:
        pred[conf < np.asarray(thresholds)[pred_idx]] = -1
    return pred, conf
