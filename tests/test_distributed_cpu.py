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

from probit_b200.distributed import exchange_unique_id, predict_sharded, restart_batch


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
        # communicator bootstrap: rank 0's 128 bytes reach every rank (NCCL's unique id travels this way)
        uid = exchange_unique_id(lambda: bytes(range(128)), device="cpu")
        ok = ok and uid == bytes(range(128))
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


# ---- the schedule of csrc/dist.cu (block-column-cyclic Cholesky + test rows carried along), modelled on CPU ----------
def _spd(n, seed):
    g = torch.Generator().manual_seed(seed)
    G = torch.randn(n, n + 3, dtype=torch.float64, generator=g)
    s = torch.rand(n, dtype=torch.float64, generator=g) + 0.5
    K = G @ G.T / n
    B = torch.eye(n, dtype=torch.float64) + s[:, None] * K * s[None, :]
    return K, s, B


def _model_worker(rank, world, port, n, nb, out):
    from bc_model import BlockCyclicModel, make_apply_hook
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K, s, B = _spd(n, 5)
        g = torch.Generator().manual_seed(7)
        m_all = 11
        Ks = torch.randn(m_all, n, dtype=torch.float64, generator=g)            # stand-in for k(X*, X)
        from probit_b200.distributed import shard_range
        lo, hi = shard_range(m_all, rank, world)                                  # test points sharded, no collective
        V = (Ks[lo:hi] * s[None, :]).clone()
        model = BlockCyclicModel(n, nb)
        logdet = [0.0]
        apply = make_apply_hook(V, n)

        def hook(k, k0, w, P):
            logdet[0] += float(torch.log(torch.diagonal(P[:w, :w])).sum())
            apply(k, k0, w, P)

        model.factor(lambda i, j: B[i, j], hook)
        L = torch.linalg.cholesky(B)
        var = 1.0 - (V * V).sum(1)
        var_ref = 1.0 - torch.einsum("ij,ij->i", Ks[lo:hi] * s, torch.cholesky_solve((Ks[lo:hi] * s).T, L).T)
        ok_var = bool(torch.allclose(var, var_ref, rtol=0, atol=1e-10))
        ok_logdet = abs(logdet[0] - float(torch.log(torch.diagonal(L)).sum())) < 1e-9
        # owned block columns hold the factor
        ok_factor = True
        for j in range(model.me, model.nblk, model.world):
            j0, w = j * nb, model.width(j)
            ok_factor = ok_factor and bool(torch.allclose(torch.tril(model.col(j)[:w]), torch.tril(L[j0:j0 + w, j0:j0 + w]), atol=1e-10))
            ok_factor = ok_factor and bool(torch.allclose(model.col(j)[w:], L[j0 + w:, j0:j0 + w], atol=1e-10))
        # a later chunk of test rows re-streams the stored panels
        V2 = (Ks[lo:hi] * s[None, :]).clone()
        model.stream(make_apply_hook(V2, n))
        ok_stream = bool(torch.allclose(V2, V, rtol=0, atol=1e-12))
        recv_order = [k for e, k in model.log if e == "recv"]
        ok_order = recv_order == list(range(model.nblk)) * 2
        out[rank] = (ok_var, ok_logdet, ok_factor, ok_stream, ok_order)
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_block_cyclic_schedule_world2_matches_lapack():
    for n, nb in [(37, 8), (64, 16), (50, 64), (96, 8)]:
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_model_worker, args=(2, _free_port(), n, nb, out), nprocs=2, join=True)
            res = dict(out)
            assert res[0] == (True,) * 5 and res[1] == (True,) * 5, (n, nb, res)


def test_block_cyclic_schedule_single_process():
    out = {}
    _model_worker(0, 1, 0, 45, 8, out)
    assert out[0] == (True,) * 5


def test_dist_layout_queries_do_not_need_a_gpu():
    """pb_dist_workspace_bytes: per-rank memory is ~2 N^2 / G (row block of K + the rank's block columns), so N = 131072
    fits eight 180 GB GPUs although one N x N matrix alone (128 GiB) plus its factor would not fit one."""
    from probit_b200 import _lib
    lib = _lib.load()
    one = lib.pb_fit_workspace_bytes(131072, 4)
    per_rank = [lib.pb_dist_workspace_bytes(131072, 4, 8, r, None) for r in range(8)]
    assert one > 250 * 2**30 and max(per_rank) < 40 * 2**30
    assert len(set(per_rank)) == 1                                  # same size on every rank (symmetric allocation)
    assert lib.pb_dist_workspace_bytes(65536, 4, 1, 0, None) < lib.pb_fit_workspace_bytes(65536, 4) + (2 << 30)
    assert lib.pb_dist_predict_scratch_bytes(65536, 4, 512) >= 512 * 65536 * 8
