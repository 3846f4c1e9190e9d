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


def warping_paths_fast(s1, s2, window=None, max_dist=None, use_pruning=False, max_step=None,
                       max_length_diff=None, penalty=None, psi=None, psi_neg=True, compact=False,
                       inner_dist="squared euclidean", **kwargs):
    """`dtaidistance.dtw.warping_paths_fast` for the one call the reference makes (sig_proc.py:298-305:
    penalty, psi = (b1, 0, b2, 0), compact=False, psi_neg=False): (distance, sqrt'ed paths matrix),
    computed by the C restatement in oracle/wdx_oracle.c (PARITY UNPINNED, see its header)."""
    from oracle import wdx_oracle as _o

    if any(v not in (None, 0) for v in (window, max_dist, max_step, max_length_diff)) or use_pruning or compact:
        raise NotImplementedError("shim supports only penalty/psi (what the reference passes)")
    if inner_dist != "squared euclidean" or psi_neg:
        raise NotImplementedError
    if psi is None:
        psi = (0, 0, 0, 0)
    elif isinstance(psi, int):
        psi = (psi, psi, psi, psi)
    elif len(psi) == 2:
        psi = (psi[0], psi[0], psi[1], psi[1])
    return _o.warping_paths(np.asarray(s1, dtype=np.float64), np.asarray(s2, dtype=np.float64), penalty or 0.0, psi)


def best_path(paths, row=None, col=None, use_max=False):
    from oracle import wdx_oracle as _o

    if use_max:
        raise NotImplementedError
    return _o.best_path(paths, row=row, col=col)
