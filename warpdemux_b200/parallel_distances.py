"""`distance_matrix_to` — drop-in for warpdemux/parallel_distances.py:48-84.

Same signature and return (float32 [nX, nY]); `block_size`, `n_jobs`, `pbar`,
`pbar_kwargs` are accepted and ignored (the reference uses them to split the
work over CPU processes).  Computed by `wdx_distance_matrix_to`
(include/wdx_b200.h) in float64 and cast on store, like the reference's
`.astype(np.float32)`."""
from typing import Optional

import numpy as np

from .device_model import distance_matrix
from .sharding import default_device


def distance_matrix_to(
    X,
    Y,
    window: Optional[int] = None,
    penalty: Optional[float] = None,
    block_size: Optional[int] = None,
    n_jobs: int = -1,
    pbar: bool = False,
    pbar_kwargs: dict = {},
    mode: str = "exact",
):
    return distance_matrix(np.asarray(X), np.asarray(Y), window, penalty, mode=mode, out_dtype=np.float32,
                           device=default_device())
