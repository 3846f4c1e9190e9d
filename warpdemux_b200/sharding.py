"""Read sharding across the GPUs of one box.

The path is embarrassingly parallel over reads (the reference exploits the
same independence with a process pool, file_proc.py:1197-1245): read i goes to
GPU g = i*G // n (contiguous index ranges), the model is replicated, and there
is NO data-path collective — only an order-preserving host-side concatenation
of the per-shard results.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def default_device() -> int:
    """LOCAL_RANK under torchrun (one process per GPU), else WDX_B200_DEVICE, else 0."""
    for var in ("WDX_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v != "":
            return int(v)
    return 0


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) per rank: read i belongs to rank i*world // n."""
    if world < 1:
        raise ValueError("world must be >= 1")
    # smallest i with i*world // n >= g  is ceil(g*n / world)
    cuts = [-(-g * n // world) for g in range(world + 1)]
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def shard_of(i: int, n: int, world: int) -> int:
    return i * world // n


def gather_in_shard_order(parts: Sequence[Sequence[np.ndarray]]) -> Tuple[np.ndarray, ...]:
    """parts[rank] = tuple of per-shard arrays -> tuple of concatenated arrays (rank order = read order)."""
    n_out = len(parts[0])
    return tuple(np.concatenate([np.asarray(p[q]) for p in parts], axis=0) for q in range(n_out))


def predict_sharded(predict_fn: Callable[[np.ndarray], Tuple[np.ndarray, ...]], X: np.ndarray,
                    group=None, x_is_local_shard: bool = False) -> Tuple[np.ndarray, ...]:
    """One process per GPU (torchrun): every rank classifies its contiguous range
    of reads with `predict_fn` (e.g. `lambda x: model.predict(x, nproc=1)`), then
    the per-rank results are gathered host-side and concatenated in rank order.
    Without an initialised process group this is just `predict_fn(X)`.

    X: the whole batch on every rank (default), or this rank's shard only
    (`x_is_local_shard=True`, e.g. when each rank read its own part of the input).
    Returns the full-length arrays on every rank."""
    try:
        import torch.distributed as dist
        active = dist.is_available() and dist.is_initialized()
    except Exception:  # noqa: BLE001
        active = False
    if not active:
        return tuple(np.asarray(a) for a in predict_fn(X))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if x_is_local_shard:
        local = X
    else:
        lo, hi = shard_bounds(len(X), world)[rank]
        local = X[lo:hi]
    mine = tuple(np.asarray(a) for a in predict_fn(local))
    parts: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(parts, mine, group=group)   # labels / probabilities: <= 100 B per read
    return gather_in_shard_order(parts)
