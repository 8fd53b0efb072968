"""Multi-GPU check of the block-cyclic Cholesky path (run under torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 4096 32768
For each N: LaplaceGP fit with the distributed factorisation vs the single-GPU factorisation on every rank
(weights must agree to 1e-10), plus potrf wall time of both.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from probit_b200 import _lib, approximators as PA, kernels as PK, utilities as PU
from probit_b200.distributed import DistributedFactorization

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for n in [int(a) for a in sys.argv[1:]]:
    rng = np.random.default_rng(3)
    X = rng.uniform(size=(n, 4))
    f = np.sin(3 * X[:, 0]) + X[:, 1] - X[:, 2] ** 2 + 0.3 * rng.standard_normal(n)
    order = np.argsort(f); y = np.empty(n, dtype=np.int64); y[order] = (np.arange(n) * 5) // n
    fs = np.sort(f); cut = np.array([-np.inf] + [0.5 * (fs[(j * n) // 5] + fs[(j * n) // 5 - 1]) for j in range(1, 5)] + [np.inf])
    params = (1.0, (float(np.sqrt(0.4)), cut))
    gp = PA.LaplaceGP((X, y), lambda l: 1.0 * PK.Matern12().stretch(l), PU.log_probit_likelihood)
    _lib.set_option("laplace_pcg_min_n", 1 << 40)      # factor every Newton step: this is a Cholesky test
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w1, p1 = gp.approximate_posterior(params)
    torch.cuda.synchronize(); t_single = time.perf_counter() - t0
    it1, f1 = gp.last_result.iterations, gp.last_result.factorizations
    with DistributedFactorization(gp) as hook:
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        w2, p2 = gp.approximate_posterior(params)
        torch.cuda.synchronize(); dist.barrier(); t_dist = time.perf_counter() - t0
        assert hook.error is None, hook.error
    err = ((w2 - w1).norm() / w1.norm()).item()
    print(f"[rank {rank}/{world}] N={n} iterations {it1}/{gp.last_result.iterations} factorizations {f1} "
          f"single-GPU fit {t_single:.3f} s  {world}-GPU fit {t_dist:.3f} s  rel diff {err:.2e}", flush=True)
    assert err < 1e-10 and it1 == gp.last_result.iterations
    # default Newton policy forced on (Nystrom-preconditioned CG): row-sharded K x + all-gather vs the local symv
    _lib.set_option("laplace_pcg_min_n", 0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w3, _ = gp.approximate_posterior(params)
    torch.cuda.synchronize(); t_cg_single = time.perf_counter() - t0
    cg1 = gp.last_result.pcg_iterations
    with DistributedFactorization(gp) as hook:
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        w4, _ = gp.approximate_posterior(params)
        torch.cuda.synchronize(); dist.barrier(); t_cg_dist = time.perf_counter() - t0
        assert hook.error is None, hook.error
        mv = hook.matvec_calls
    err_cg = ((w4 - w3).norm() / w3.norm()).item()
    print(f"[rank {rank}/{world}] N={n} CG Newton: single {t_cg_single:.3f} s ({cg1} CG)  sharded matvec {t_cg_dist:.3f} s "
          f"({gp.last_result.pcg_iterations} CG, {mv} sharded products)  rel diff {err_cg:.2e}  vs factor path "
          f"{((w3 - w1).norm() / w1.norm()).item():.2e}", flush=True)
    assert err_cg < 1e-10 and mv > 0
    _lib.set_option("laplace_pcg_min_n", 24576)
    del gp
    torch.cuda.empty_cache()
dist.destroy_process_group()
