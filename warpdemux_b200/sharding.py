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


def _worker_index() -> Optional[int]:
    """0-based index of this process inside a multiprocessing / concurrent.futures pool, None in a main process.
    The reference runs its minibatches in a ProcessPoolExecutor whose workers all inherit ONE environment
    (file_proc.py:1197-1245), so the pool index is the only thing that tells the workers apart."""
    try:
        import multiprocessing

        ident = multiprocessing.current_process()._identity
        return int(ident[-1]) - 1 if ident else None
    except Exception:  # noqa: BLE001
        return None


def visible_devices() -> List[int]:
    """Devices the workers of this process tree may use: WDX_B200_DEVICES="0,2,5" or "all" (default: all CUDA devices)."""
    v = os.environ.get("WDX_B200_DEVICES", "all").strip()
    if v and v != "all":
        return [int(t) for t in v.split(",") if t.strip() != ""]
    try:
        from . import _lib

        return list(range(max(1, _lib.device_count())))
    except Exception:  # noqa: BLE001
        return [0]


def default_device() -> int:
    """Which GPU a handle created without an explicit device lands on:
      1. WDX_B200_DEVICE (explicit), 2. LOCAL_RANK (torchrun: one process per GPU),
      3. pool workers (the reference's ProcessPoolExecutor): round-robin over `visible_devices()` by pool index, so
         `num_proc` workers spread over every GPU of the box without any per-worker configuration,
      4. else the first visible device."""
    for var in ("WDX_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v != "":
            return int(v)
    devs = visible_devices()
    w = _worker_index()
    return devs[w % len(devs)] if w is not None else devs[0]


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


def _gather_arrays(mine: Tuple[np.ndarray, ...], group, dst: Optional[int]) -> Optional[List[tuple]]:
    """Per-rank tuples of arrays -> list over ranks (on every rank, or on `dst` only; None elsewhere).  Arrays travel as
    raw tensors (shards differ by at most one read: padded to the longest), not as pickles: 12.5 M labels per rank are
    100 MB, which `all_gather_object` would serialise byte by byte."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta: List[Optional[list]] = [None] * world
    dist.all_gather_object(meta, [(a.shape, a.dtype.str) for a in mine], group=group)    # a few bytes
    out: Optional[List[list]] = [[] for _ in range(world)] if (dst is None or rank == dst) else None
    for q, a in enumerate(mine):
        shapes = [m[q][0] for m in meta]
        dt = np.dtype(meta[0][q][1])
        row = int(np.prod(shapes[0][1:])) if len(shapes[0]) > 1 else 1
        longest = max(int(sh[0]) for sh in shapes) * row * dt.itemsize
        buf = np.zeros(longest, dtype=np.uint8)
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        buf[: raw.size] = raw
        t = torch.from_numpy(buf).to(dev)
        if dst is None:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t, group=group)
        else:
            parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
            dist.gather(t, parts, dst=dst, group=group)
        if out is not None:
            for r in range(world):
                nbytes = int(np.prod(shapes[r])) * dt.itemsize
                out[r].append(parts[r][:nbytes].cpu().numpy().view(dt).reshape(shapes[r]))
    return [tuple(p) for p in out] if out is not None else None


def predict_sharded(predict_fn: Callable[[np.ndarray], Tuple[np.ndarray, ...]], X: np.ndarray,
                    group=None, x_is_local_shard: bool = False, dst: Optional[int] = None):
    """One process per GPU (torchrun): every rank classifies its contiguous range
    of reads with `predict_fn` (e.g. `lambda x: model.predict(x, nproc=1)`), then
    the per-rank results are gathered host-side and concatenated in rank order.
    Without an initialised process group this is just `predict_fn(X)`.

    X: the whole batch on every rank (default), or this rank's shard only
    (`x_is_local_shard=True`, e.g. when each rank read its own part of the input).
    Returns the full-length arrays on every rank; with `dst=r` only rank r receives them (the others get None) —
    the host-side label gather of the reference's main process."""
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
    parts = _gather_arrays(mine, group, dst)            # labels / probabilities: <= 100 B per read
    return gather_in_shard_order(parts) if parts is not None else None
