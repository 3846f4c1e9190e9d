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


# -- the block-parallel variant of the reference (parallel_distances.py:24-45, 87-198) -------------
# The reference cuts the matrix into block_size x block_size tiles over a process pool; on the GPU the whole
# (sub)matrix is one launch, so `block_size` / `n_jobs` / `pbar*` are accepted and ignored.  Same float32 results.
def compute_block_distance(block_indices, X, window=None, penalty=None, mode: str = "exact", **kwargs):
    """(i, j, float32 [len(i), len(j)]) for the rows X[i] against X[j] (parallel_distances.py:24-45)."""
    i, j = block_indices
    i, j = np.asarray(i), np.asarray(j)
    X = np.asarray(X)
    return i, j, distance_matrix_to(X[i], X[j], window=window, penalty=penalty, mode=mode)


def parallel_distance_matrix(X, block_size: int = 1000, n_jobs: int = 6, subset=None, window: Optional[int] = None,
                             penalty: Optional[float] = None, pbar: bool = False, pbar_kwargs: dict = {},
                             mode: str = "exact", **kwargs):
    """float32 [r1_end - r1_start, r2_end - r2_start] (parallel_distances.py:139-198); `subset` =
    ((r1_start, r1_end), (r2_start, r2_end)), default: all rows against all rows."""
    X = np.asarray(X)
    if subset:
        (r1_start, r1_end), (r2_start, r2_end) = subset
    else:
        r1_start, r1_end, r2_start, r2_end = 0, X.shape[0], 0, X.shape[0]
    return distance_matrix_to(X[r1_start:r1_end], X[r2_start:r2_end], window=window, penalty=penalty, mode=mode)


def parallel_distance_matrix_to(X, Y, block_size: int = 1000, n_jobs: int = 6, window: Optional[int] = None,
                                penalty: Optional[float] = None, pbar: bool = False, pbar_kwargs: dict = {},
                                mode: str = "exact", **kwargs):
    """float32 [nX, nY] (parallel_distances.py:87-136)."""
    return distance_matrix_to(X, Y, window=window, penalty=penalty, mode=mode)
