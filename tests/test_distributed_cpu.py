"""world_size-2 gloo tests of the multi-GPU host logic (probit_b200/distributed.py) on CPU.

The compute function is injected (a NumPy stand-in), so what is tested is the partitioning, the
padding of ragged shards and the gather order — the parts that run above the C ABI.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probit_b200.distributed import predict_sharded, restart_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_test, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = torch.arange(n_test * 2, dtype=torch.float64).reshape(n_test, 2)
        fake_predict = lambda Xs: (Xs.sum(1), (Xs * Xs).sum(1))
        m, v, rng = predict_sharded(fake_predict, X, gather=True)
        ok = torch.equal(m, X.sum(1)) and torch.equal(v, (X * X).sum(1)) and rng == (0, n_test)
        ms, vs, (lo, hi) = predict_sharded(fake_predict, X, gather=False)
        ok = ok and torch.equal(ms, X[lo:hi].sum(1)) and ms.numel() in (n_test // world, n_test // world + 1)
        params = [float(i) for i in range(5)]
        seen = []
        def evaluate(p):
            seen.append(p)
            return [p * p, -p]
        full = restart_batch(evaluate, params)
        ok = ok and seen == params[rank::world]
        ok = ok and torch.equal(full, torch.tensor([[p * p, -p] for p in params], dtype=torch.float64))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_predict_sharding_and_restart_batch_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), 11, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_single_process_paths():
    X = torch.arange(10, dtype=torch.float64).reshape(5, 2)
    m, v, rng = predict_sharded(lambda Xs: (Xs.sum(1), Xs.prod(1)), X, gather=True)
    assert rng == (0, 5) and torch.equal(m, X.sum(1))
    full = restart_batch(lambda p: p + 1.0, [1.0, 2.0, 3.0])
    assert torch.equal(full, torch.tensor([[2.0], [3.0], [4.0]], dtype=torch.float64))
