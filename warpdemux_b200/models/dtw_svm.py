"""`DTW_SVM` — drop-in for the reference's class of the same name
(warpdemux/models/dtw_svm.py:25-98).  `predict` has the reference's signature,
return types and error behaviour; the work happens in ONE call into the CUDA
library (distance matrix, kernel, libsvm probabilities and thresholding fused;
include/wdx_b200.h `wdx_predict`)."""
import logging
import os
from typing import Optional, Tuple, Union

import numpy as np
import pandas as pd

from .. import model_io
from .dtw_base import BaseDTWModel
from .utils import predictions_to_df

#: arithmetic mode of the DTW recurrence: "exact" (float64, bit-exact),
#: "fast" (float32), "guarded" (float32 + exact recompute near decision boundaries)
DEFAULT_MODE = os.environ.get("WDX_B200_MODE", "guarded")


class DTW_SVM(BaseDTWModel):
    def __init__(self, params: Optional[model_io.ModelParams] = None, device: Optional[int] = None,
                 mode: str = DEFAULT_MODE, on_nonfinite: str = "raise"):
        super().__init__(params=params, device=device)
        if mode not in ("exact", "fast", "guarded"):
            raise ValueError(f"mode must be exact|fast|guarded, got {mode!r}")
        if on_nonfinite not in ("raise", "noise"):
            raise ValueError("on_nonfinite must be 'raise' or 'noise'")
        self.mode = mode
        self.on_nonfinite = on_nonfinite

    # the reference's hyper-parameter attributes (dtw_svm.py:26-30)
    @property
    def gamma(self):
        return self.params.gamma

    @property
    def pwr_dist(self):
        return self.params.pwr_dist

    # -- construction ---------------------------------------------------------
    @classmethod
    def from_reference(cls, ref_model, **kw) -> "DTW_SVM":
        """From an unpickled reference `DTW_SVM` object."""
        return cls(model_io.from_reference_model(ref_model, name=getattr(ref_model, "name", "")), **kw)

    @classmethod
    def load(cls, path: str, **kw) -> "DTW_SVM":
        """From a reference `.joblib` model file or this package's `.npz`."""
        return cls(model_io.load_model(path), **kw)

    # -- the seam -------------------------------------------------------------
    def predict(
        self,
        X: np.ndarray,
        nproc: int = -1,
        block_size: Optional[int] = None,
        pbar: bool = False,
        pbar_kwargs: dict = {},
        return_df: bool = False,
    ) -> Union[Tuple[np.ndarray, np.ndarray], pd.DataFrame]:
        """`nproc`, `block_size`, `pbar`, `pbar_kwargs` are accepted for
        signature compatibility and ignored: the GPU does the whole batch."""
        if not self.is_trained:
            msg = "Model not trained yet."
            logging.error(msg)
            raise ValueError(msg)

        X = np.asarray(X)
        if X.ndim == 1:
            X = X.reshape(1, -1)

        if X.shape[1] != self._X.shape[1]:
            raise ValueError(
                "X must have the same number of columns as the training data "
                f" ({self._X.shape[1]})."
            )

        y_pred, y_prob, conf, flags = self._device_model().predict(X, mode=self.mode)

        if flags.any() and (flags & 1).any() and self.on_nonfinite == "raise":
            # sklearn's input validation in SVC.predict_proba rejects a NaN kernel
            # matrix (reference call site dtw_svm.py:92); same exception type.
            raise ValueError("Input contains NaN.")

        if return_df:
            return predictions_to_df(y_pred, y_prob, conf, self.label_mapper)
        return y_pred, y_prob
