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

from probit_b200.distributed import predict_sharded, restart_batch, row_shard, sharded_matvec


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


# ---- block-cyclic Cholesky host logic (numpy/torch-CPU ops injected) ----------------------------------
class _CpuOps:
    def empty(self, rows, cols):
        return torch.zeros((rows, cols), dtype=torch.float64)

    def potrf_panel(self, blk, w):
        d = blk[:w, :w]
        Lc = torch.linalg.cholesky(torch.tril(d) + torch.tril(d, -1).T)
        blk[:w, :w] = torch.tril(Lc) + torch.triu(d, 1)          # strict upper left as it was
        if blk.shape[0] > w:
            blk[w:, :] = torch.linalg.solve_triangular(Lc, blk[w:, :].T, upper=False).T
        return torch.zeros(1, dtype=torch.int32)

    def gemm_nt(self, A, B, C_out, alpha, beta):
        C_out.copy_(alpha * (A @ B.T) + beta * C_out)


def _chol_worker(rank, world, port, n, nb, out):
    from probit_b200.distributed import BlockCyclicCholesky
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        G = torch.randn(n, n + 3, dtype=torch.float64, generator=g)
        A = G @ G.T + n * torch.eye(n, dtype=torch.float64)
        Lfull = torch.zeros(n, n, dtype=torch.float64)
        chol = BlockCyclicCholesky(n, _CpuOps(), nb=nb)
        seen = []

        def fill(j0, w, o):
            o.copy_(A[j0:, j0:j0 + w])

        def write(k0, w, panel):
            seen.append(k0)
            Lfull[k0:, k0:k0 + w] = panel

        chol.factor(fill, write)
        ref = torch.linalg.cholesky(A)
        err = (torch.tril(Lfull) - ref).abs().max().item()
        out[rank] = (err < 1e-10, seen == [k * nb for k in range(chol.nblk)], len(chol.owned))
    finally:
        dist.destroy_process_group()


def test_block_cyclic_cholesky_world2_matches_lapack():
    world = 2
    for n, nb in [(37, 8), (64, 16), (50, 64)]:
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_chol_worker, args=(world, _free_port(), n, nb, out), nprocs=world, join=True)
            res = dict(out)
            assert res[0][:2] == (True, True) and res[1][:2] == (True, True), (n, nb, res)
            assert res[0][2] + res[1][2] == (n + nb - 1) // nb


def test_block_cyclic_cholesky_single_process():
    from probit_b200.distributed import BlockCyclicCholesky
    n, nb = 45, 8
    G = torch.randn(n, n, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    A = G @ G.T + n * torch.eye(n, dtype=torch.float64)
    Lfull = torch.zeros(n, n, dtype=torch.float64)
    chol = BlockCyclicCholesky(n, _CpuOps(), nb=nb)
    chol.factor(lambda j0, w, o: o.copy_(A[j0:, j0:j0 + w]), lambda k0, w, p: Lfull[k0:, k0:k0 + w].copy_(p))
    assert (torch.tril(Lfull) - torch.linalg.cholesky(A)).abs().max().item() < 1e-10


def _matvec_worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        K = torch.randn(n, n, dtype=torch.float64, generator=g)
        K = K + K.T
        buffers, ok = {}, True
        for trial in range(3):                                   # buffers are reused between calls
            x = torch.randn(n, dtype=torch.float64, generator=g)
            seen = []
            def local_product(lo, hi, dst):
                seen.append((lo, hi))
                dst[: hi - lo] = K[lo:hi] @ x
            y = sharded_matvec(local_product, n, buffers=buffers)
            lo, hi, chunk = row_shard(n, rank, world)
            ok = ok and seen == ([(lo, hi)] if hi > lo else []) and y.numel() == n
            ok = ok and torch.allclose(y, K @ x, rtol=0, atol=1e-12)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_sharded_matvec_world2_ragged():
    for n in (7, 64):                                            # 7: ragged last shard
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_matvec_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
            assert dict(out) == {0: True, 1: True}


def test_row_shard_covers_every_row_once():
    for n in (1, 5, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [row_shard(n, r, world) for r in range(world)]
            rows = [i for lo, hi, _ in spans for i in range(lo, hi)]
            assert rows == list(range(n)) and len({c for _, _, c in spans}) == 1
