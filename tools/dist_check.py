"""Multi-GPU check of the partitioned path (run under torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_check.py 4096 32768 [--single-max 65536] [--test-n 2048] [--json out.json]

For each N: ONE ShardedLaplaceGP fit + predict over all the ranks (pb_dist_laplace_fit / pb_dist_predict: rows of K
sharded, NCCL collectives enqueued from C++, block-column-cyclic factor, test points sharded) against the single-GPU
LaplaceGP run on every rank (skipped above --single-max, where one GPU cannot hold K and the factor).  Weights,
precisions, predictive moments and the objective must agree to 1e-10, iteration counts must be equal.  Above
--single-max the fixed-point residual ||grad_ll(K w) - w|| of the returned weight is checked instead (the same
size-independent property tests/test_gpu_fit.py uses at N = 65536).  Prints per-stage wall times.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from probit_b200 import approximators as PA, kernels as PK, utilities as PU  # noqa: E402
from probit_b200.distributed import ShardedLaplaceGP, shard_range  # noqa: E402


def problem(n):
    rng = np.random.default_rng(3)
    X = rng.uniform(size=(n, 4))
    f = np.sin(3 * X[:, 0]) + X[:, 1] - X[:, 2] ** 2 + 0.3 * rng.standard_normal(n)
    order = np.argsort(f)
    y = np.empty(n, dtype=np.int64)
    y[order] = (np.arange(n) * 5) // n
    fs = np.sort(f)
    cut = np.array([-np.inf] + [0.5 * (fs[(j * n) // 5] + fs[(j * n) // 5 - 1]) for j in range(1, 5)] + [np.inf])
    return X, y, (1.0, (float(np.sqrt(0.4)), cut))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sizes", type=int, nargs="+")
    ap.add_argument("--single-max", type=int, default=65536)
    ap.add_argument("--test-n", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    prior = lambda l: 1.0 * PK.Matern12().stretch(l)  # noqa: E731
    records = []

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for n in args.sizes:
        X, y, params = problem(n)
        Xs = np.random.default_rng(5).uniform(-0.2, 1.2, size=(args.test_n, 4))
        lo, hi = shard_range(args.test_n, rank, world)
        rec = {"n": n, "world": world, "n_test": args.test_n}
        gp = ShardedLaplaceGP((X, y), prior, PU.log_probit_likelihood)
        Xs_loc = torch.as_tensor(Xs[lo:hi], device="cuda")
        for rep in range(args.reps):
            sync(); t0 = time.perf_counter()
            w, p = gp.approximate_posterior(params)
            sync(); t1 = time.perf_counter()
            m, v = gp.predict(Xs_loc, params, w, p)
            sync(); t2 = time.perf_counter()
        rec.update(fit_s=t1 - t0, predict_s=t2 - t1, iterations=gp.last_result.iterations,
                   cg_iterations=gp.last_result.pcg_iterations)
        sync(); t0 = time.perf_counter()
        obj = gp.objective()(params)
        sync(); rec["objective_s"] = time.perf_counter() - t0
        rec["ws_gib"] = gp._ws_bytes / 2 ** 30
        # size-independent property: the returned weight is a fixed point of w -> grad_ll(K w)
        f = gp.posterior_mean(w, params)
        g = PU.evaluate_likelihood(gp._kind, f, gp.y, params[1], ("g",), gp.likelihood_eps)["g"]
        rec["fixed_point_residual"] = float((g - w).norm() / w.norm())
        assert rec["fixed_point_residual"] < 1e-6, rec
        assert bool(torch.isfinite(v).all()) and float(v.min()) > 0 and float(v.max()) <= 1.0 + 1e-9
        if n <= args.single_max:
            del gp
            torch.cuda.empty_cache()
            # the same Newton policy (Nystrom-preconditioned CG) on one GPU, so that only the partitioning differs
            ref = PA.LaplaceGP((X, y), prior, PU.log_probit_likelihood, options=dict(laplace_pcg_min_n=0))
            for rep in range(args.reps):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                w1, p1 = ref.approximate_posterior(params)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                m1, v1 = ref.predict(Xs_loc, params, w1, p1)
                torch.cuda.synchronize(); t2 = time.perf_counter()
            obj1 = ref.objective()(params)
            rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
            rec.update(single_fit_s=t1 - t0, single_predict_s=t2 - t1, single_iterations=ref.last_result.iterations,
                       single_cg_iterations=ref.last_result.pcg_iterations,
                       err_weight=rel(w, w1), err_precision=rel(p, p1), err_mean=rel(m, m1), err_variance=rel(v, v1),
                       err_objective=abs(obj - obj1) / abs(obj1))
            del ref
            torch.cuda.empty_cache()
            assert rec["iterations"] == rec["single_iterations"], rec
            assert max(rec["err_weight"], rec["err_precision"], rec["err_mean"], rec["err_variance"], rec["err_objective"]) < 1e-10, rec
        else:
            del gp
            torch.cuda.empty_cache()
        print(f"[rank {rank}/{world}] " + json.dumps(rec), flush=True)
        records.append(rec)
    if rank == 0 and args.json:
        with open(args.json, "w") as fh:
            json.dump(records, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
