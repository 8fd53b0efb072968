"""bench.py — the hot path of BASELINE.json on B200: LaplaceGP ordinal (J=5) approximate_posterior + predict.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full pass of the hot path over the synthetic C4 workload (BASELINE.json configs[3]):
N=65536, D=4, J=5, Matern12 (l=1, s^2=1), sigma=sqrt(0.4), tol=1e-5, FP64: Gram -> Newton to
convergence (one Cholesky per iteration) -> precision -> predict (one more Cholesky + N_test-RHS solve)
at N_test=4096.  `value` times K steps with all inputs resident in HBM; `e2e` times K steps through the
public class API from pinned HOST buffers (H2D of X, y, X_test and D2H of weight, precision, mean,
variance inside the timed region).  Multi-GPU (N>1): one hyperparameter restart per GPU (BASELINE
north_star: "independent hyperparameter or restart batches map one per GPU"), no data-path collective,
weak scaling; value = max-over-ranks step time / N (seconds per fit+predict at aggregate throughput).

The reference arm (--impl reference) and the cpu_baseline leg time the NumPy/SciPy oracle port in the
reference's literal operation sequence (dense Jacobian + LU, JAX itself is not installable here) on a
bounded sample (smaller N) and scale by N^3 — see `sample` in the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "approximate_posterior+predict seconds at N=64k FP64"
SEED = 1
NOISE_VARIANCE = 0.4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-n", dest="n", type=int, default=65536, help="training points (65536 = BASELINE config; smaller only for debugging)")
    ap.add_argument("--test-n", dest="n_test", type=int, default=4096)
    ap.add_argument("--cpu-sample-n", dest="cpu_n", type=int, default=4096, help="bounded-sample size of the CPU baseline / reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="restarts", choices=["restarts", "cholesky"],
                    help="multi-GPU mode: one hyperparameter restart per GPU (weak scaling, default) or ONE fit whose "
                         "Cholesky factorisations are block-cyclic across the GPUs and whose test points are sharded "
                         "(strong scaling)")
    return ap.parse_args()


def workload_name(n, n_test):
    return (f"synthetic ordinal probit J=5, N={n}, D=4, Matern12 l=1 s2=1, sigma=sqrt(0.4), LaplaceGP Newton to "
            f"convergence tol=1e-5 + predict N_test={n_test}, FP64 (BASELINE configs[3])")


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_problem(n, n_test):
    from oracle import kernels as OK
    from probit_b200.datasets import generate_ordinal_data

    def sampler(X, z):
        K = (1.0 * OK.Matern12().stretch(1.0))(X) + 1e-6 * np.eye(len(X))
        return np.linalg.cholesky(K) @ z

    X, g, y, cut = generate_ordinal_data(SEED, n, 4, 5, NOISE_VARIANCE, sampler)
    Xs = np.random.default_rng(SEED + 1).uniform(-0.5, 1.5, size=(n_test, 4))
    return X, y, cut, Xs


def cpu_step(X, y, cut, Xs):
    from oracle import approximators as OA, kernels as OK, utilities as OU
    gp = OA.LaplaceGP((X, y), lambda l: 1.0 * OK.Matern12().stretch(l), OU.log_probit_likelihood,
                      newton_form="lu_jacobian")
    params = (1.0, (float(np.sqrt(NOISE_VARIANCE)), cut))
    w, p = gp.approximate_posterior(params)
    m, v = gp.predict(Xs, params, w, p)
    return len(gp.trace), float(m[0] + v[0])


def cpu_sample_info(args, seconds):
    scale = (args.n / args.cpu_n) ** 3
    return {
        "kind": "port",
        "cores": os.cpu_count(),
        "sample": (f"NumPy/SciPy oracle port, literal reference sequence (dense Jacobian + LU solve per Newton "
                   f"iteration, LU predict) at N={args.cpu_n}, N_test={min(args.n_test, args.cpu_n)}: "
                   f"{seconds:.2f} s measured, scaled by (N/{args.cpu_n})^3 = {scale:.0f}x to N={args.n} "
                   f"(EXTRAPOLATED: the literal path needs 3 NxN buffers = 96 GiB at N=65536); JAX is not installable here"),
        "measured_seconds": seconds,
        "value": seconds * scale,
        "unit": "s",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, cut, Xs = cpu_problem(args.cpu_n, min(args.n_test, args.cpu_n))
    for _ in range(args.warmup):
        cpu_step(X, y, cut, Xs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(X, y, cut, Xs)
    per_step = (time.perf_counter() - t0) / max(args.steps, 1)
    info = cpu_sample_info(args, per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": info["value"], "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["value"] * 1e3, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.n_test), "parallelism": "host cores (OpenBLAS threads)"},
        "cpu_baseline": info,
        "e2e": {"value": info["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.file.read().strip().splitlines():
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from probit_b200 import _lib, approximators as PA, kernels as PK, utilities as PU
    from probit_b200.datasets import device_latent_sampler, generate_ordinal_data

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: probit_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()

    n, n_test = args.n, args.n_test
    gen_kernel = 1.0 * PK.Matern12().stretch(1.0)                   # examples/classification.py:375
    X, g, y, cut = generate_ordinal_data(SEED, n, 4, 5, NOISE_VARIANCE, device_latent_sampler(gen_kernel, 1e-6))
    Xs = np.random.default_rng(SEED + 1).uniform(-0.5, 1.5, size=(n_test, 4))
    torch.cuda.empty_cache()
    # restart batch: rank r evaluates lengthscale l_r (geometric spread around the generating value 1.0)
    cholesky_mode = world > 1 and args.parallelism == "cholesky"
    lengthscale = 1.0 if cholesky_mode else float(2.0 ** ((rank - (world - 1) / 2.0) / 8.0))
    params = (lengthscale, (float(np.sqrt(NOISE_VARIANCE)), cut))
    prior = lambda l: 1.0 * PK.Matern12().stretch(l)

    X_pin = torch.from_numpy(X).pin_memory()
    y_pin = torch.from_numpy(y).pin_memory()
    Xs_pin = torch.from_numpy(Xs).pin_memory()
    Xd, yd, Xsd = X_pin.cuda(), y_pin.cuda(), Xs_pin.cuda()

    gp = PA.LaplaceGP((Xd, yd), prior, PU.log_probit_likelihood, tolerance=1e-5)
    hook = None
    if cholesky_mode:
        from probit_b200.distributed import DistributedFactorization, shard_range
        hook = DistributedFactorization(gp)
        hook.__enter__()
        lo, hi = shard_range(n_test, rank, world)          # test points sharded, no collective on the data path
        Xsd = Xsd[lo:hi].contiguous()
        Xs_pin = Xs_pin[lo:hi].contiguous().pin_memory()

    def step_resident():
        w, p = gp.approximate_posterior(params)
        m, v = gp.predict(Xsd, params, w, p)
        return w, p, m, v

    def step_e2e():
        # the public API from host buffers: H2D of the inputs, fit, predict, D2H of the results
        gp.X.copy_(X_pin, non_blocking=True)
        gp.y.copy_(y_pin, non_blocking=True)
        xs = Xs_pin.cuda(non_blocking=True)
        w, p = gp.approximate_posterior(params)
        m, v = gp.predict(xs, params, w, p)
        return w.cpu(), p.cpu(), m.cpu(), v.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.pb_launch_count()
    lib.pb_profile_begin()
    ms_res, out = timed(step_resident, args.steps)
    n_l, g_ms, g_fl = C.c_longlong(0), C.c_double(0), C.c_double(0)
    lib.pb_profile_end(C.byref(n_l), C.byref(g_ms), C.byref(g_fl))
    launches = lib.pb_launch_count() - launches0
    clocks = sampler.stop()
    iterations = gp.last_result.iterations
    fit_factorizations = gp.last_result.factorizations
    pcg_iterations = gp.last_result.pcg_iterations
    ms_e2e, out_h = timed(step_e2e, args.steps)

    h2d = X_pin.numel() * 8 + y_pin.numel() * 8 + Xs_pin.numel() * 8
    d2h = sum(t.numel() * 8 for t in out_h)
    sec_res = ms_res / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    if hook is not None:
        hook.__exit__(None, None, None)
    units = 1 if cholesky_mode else world      # fits completed per step across the job

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = 37.0, "fallback 37.0 TFLOP/s (tools/fp64_peak.cu, earlier run)"
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")) as fh:
            peak = float(json.load(fh)["dmma_m16n8k8_tflops_w32"])
            peak_src = ("FP64 DMMA register-resident mma.sync loop measured on this pool's B200 by tools/fp64_peak.cu "
                        "(profiles/r01_fp64_peaks.json); MEASURED_PEAKS.json carries no FP64 figure")
    except (OSError, KeyError, ValueError):
        pass
    achieved = g_fl.value / g_ms.value * 1e-9 if g_ms.value > 0 else None
    n_potrf = fit_factorizations + 1           # + the factorisation of B(w*) that predict needs
    line = {
        "metric": METRIC, "value": sec_res / units, "unit": "s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": False,
        "scaling": "strong" if cholesky_mode else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(n, n_test),
            "parallelism": ("single GPU" if world == 1 else
                            (f"one fit: block-column-cyclic Cholesky over {world} GPUs (NCCL panel broadcasts), test points sharded"
                             if cholesky_mode else
                             f"restarts: one hyperparameter restart per GPU x{world}, no data-path collective")),
            "newton_iterations": iterations, "cholesky_per_step": n_potrf, "pcg_iterations_per_step": pcg_iterations,
            "newton_policy": ("Nystrom-preconditioned CG (no factorisation)" if fit_factorizations == 0 else
                              "Cholesky of B, then PCG on the stale factor" if pcg_iterations else "Cholesky of B every step"),
            "l2": "inputs larger than L2 (K and the factor are 32 GiB each)",
            "data_generator": "classification.py:181-322 recipe, numpy default_rng(1); latent draw by the product's own Gram + potrf",
        },
        "e2e": {"value": sec_e2e / units, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": "gemm_nt_kernel<128x64, 8 warps, 2 CTA/SM> (Cholesky trailing update / TRSM / predict solve)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
            "traffic": 3.03e9, "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of ONE profiled launch of this kernel "
                                                "(SYRK 16384 x 512, algorithmic 2.21e9 B; profiles/r01_gemm_main_ncu.md); the kernel is "
                                                "tensor-bound, launches in the timed region vary in shape"),
            "launches": int(n_l.value), "kernel_ms_total": g_ms.value, "peak_source": peak_src,
            "algorithmic_flops_per_step": g_fl.value / args.steps,
        },
        "cholesky": {"n": n, "per_step": n_potrf, "flops_each": n ** 3 / 3.0,
                     "tflops_lower_bound": n_potrf * n ** 3 / 3.0 / sec_res * 1e-12},
    }
    if not args.no_cpu_baseline and world == 1:
        Xc, yc, cutc, Xsc = cpu_problem(args.cpu_n, min(n_test, args.cpu_n))
        t0 = time.perf_counter()
        cpu_step(Xc, yc, cutc, Xsc)
        line["cpu_baseline"] = cpu_sample_info(args, time.perf_counter() - t0)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
