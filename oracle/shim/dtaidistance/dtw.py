"""`dtaidistance.dtw.distance_matrix` with the upstream calling/return
convention the reference relies on (parallel_distances.py:34-43, 59-67):
a full (n x n) float64 matrix, +inf everywhere except the requested block
(upper triangle only), computed by the C restatement in oracle/wdx_oracle.c."""
import numpy as np


def distance_matrix(s, max_dist=None, use_pruning=False, max_length_diff=None, window=None,
                    max_step=None, penalty=None, psi=None, block=None, compact=False,
                    parallel=False, use_c=False, use_mp=False, show_progress=False,
                    only_triu=False, inner_dist="squared euclidean"):
    from oracle import wdx_oracle as _o

    if any(v not in (None, 0) for v in (max_dist, max_length_diff, max_step, psi)) or use_pruning:
        raise NotImplementedError("shim supports only window/penalty (what the reference passes)")
    if inner_dist != "squared euclidean" or compact:
        raise NotImplementedError
    s = np.ascontiguousarray(np.asarray(s, dtype=np.float64))
    n = s.shape[0]
    out = np.full((n, n), np.inf, dtype=np.float64)
    if block is None:
        block = ((0, n), (0, n))
    (rb, re), (cb, ce) = block[0], block[1]
    for r in range(rb, re):
        c0 = max(r + 1, cb)
        if c0 >= ce:
            continue
        out[r, c0:ce] = _o.dtw_matrix(s[r:r + 1], s[c0:ce], window or 0, penalty or 0.0)[0]
    if not only_triu:
        iu = np.triu_indices(n, 1)
        out.T[iu] = out[iu]
    return out


def warping_paths_fast(*a, **k):  # imported by sig_proc.py:16 (tRNA path only)
    raise NotImplementedError("dtaidistance shim: warping_paths_fast not restated")
