"""Read sharding across the GPUs of one box.

The path is embarrassingly parallel over reads (the reference exploits the
same independence with a process pool, file_proc.py:1197-1245): read i goes to
GPU g = i*G // n (contiguous index ranges), the model is replicated, and there
is NO data-path collective — only an order-preserving host-side concatenation
of the per-shard labels.
"""
from __future__ import annotations

import os
from typing import List, Tuple


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
