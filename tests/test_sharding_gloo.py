"""N>1 host logic on CPU: two processes (gloo), contiguous read shards, host-side
gather in shard order.  The per-shard classifier is the CPU oracle here (no GPU
in this container); on the GPU box the same `predict_sharded` wraps DTW_SVM.predict."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle import wdx_oracle as o
    from warpdemux_b200 import model_io
    from warpdemux_b200.sharding import predict_sharded, shard_bounds
    from wdx_testutil import synth_fingerprints

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", "WDX4_rna004_v1_0.npz"))
    X = synth_fingerprints(m.sv, n, seed=5)
    calls = []

    def fn(x):
        calls.append(len(x))
        pred, prob, conf, _ = o.predict(m, x) if len(x) else (np.zeros(0, np.int64), np.zeros((0, m.k)), np.zeros(0), None)
        return pred, prob, conf

    full = predict_sharded(fn, X)
    lo, hi = shard_bounds(n, world)[rank]
    local = predict_sharded(fn, X[lo:hi], x_is_local_shard=True)
    root_only = predict_sharded(fn, X, dst=0)              # the main process's label gather: rank 0 alone receives
    assert (root_only is None) == (rank != 0)
    if rank == 0:
        assert all(np.array_equal(a, b) for a, b in zip(root_only, full))
    q.put((rank, calls, full[0], full[1], local[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 1])
def test_two_rank_shards_gather_in_read_order(n, models):
    import torch.multiprocessing as mp

    from oracle import wdx_oracle as o
    from wdx_testutil import synth_fingerprints

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = models["WDX4_rna004_v1_0"]
    want_pred, want_prob, _, _ = o.predict(m, synth_fingerprints(m.sv, n, seed=5))
    sizes = [out[0][1][0], out[1][1][0]]
    assert sum(sizes) == n and abs(sizes[0] - sizes[1]) <= 1      # each rank saw only its contiguous range
    for rank, calls, pred, prob, local_pred in out:
        assert np.array_equal(pred, want_pred) and np.array_equal(prob, want_prob)   # full result on every rank
        assert np.array_equal(local_pred, want_pred)


def test_predict_sharded_without_process_group(models):
    from warpdemux_b200.sharding import gather_in_shard_order, predict_sharded

    X = np.arange(10.0).reshape(5, 2)
    out = predict_sharded(lambda x: (x[:, 0] * 2, x), X)
    assert np.array_equal(out[0], X[:, 0] * 2)
    parts = [(np.array([1, 2]), np.zeros((2, 3))), (np.array([3]), np.ones((1, 3)))]
    a, b = gather_in_shard_order(parts)
    assert a.tolist() == [1, 2, 3] and b.shape == (3, 3)


def _which_device(_):
    from warpdemux_b200.sharding import default_device

    return default_device()


def test_pool_workers_spread_over_the_visible_devices(monkeypatch):
    """The reference's ProcessPoolExecutor forks its workers with ONE environment (file_proc.py:1197-1245): handles
    created without an explicit device are assigned round-robin by pool index; explicit settings win."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    from warpdemux_b200 import sharding

    for var in ("WDX_B200_DEVICE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    monkeypatch.setenv("WDX_B200_DEVICES", "0,1,2,3")
    assert sharding.default_device() == 0 and sharding.visible_devices() == [0, 1, 2, 3]
    with ProcessPoolExecutor(8, mp_context=mp.get_context("fork")) as ex:
        got = sorted(set(ex.map(_which_device, range(64), chunksize=1)))
    assert set(got) <= {0, 1, 2, 3} and len(got) >= 2          # more than one device is in use ...
    monkeypatch.setenv("WDX_B200_DEVICES", "2,5")
    with ProcessPoolExecutor(4, mp_context=mp.get_context("fork")) as ex:
        assert set(ex.map(_which_device, range(32), chunksize=1)) <= {2, 5}
    monkeypatch.setenv("LOCAL_RANK", "3")                        # ... unless torchrun / the user pinned it
    assert sharding.default_device() == 3
    monkeypatch.setenv("WDX_B200_DEVICE", "1")
    assert sharding.default_device() == 1
