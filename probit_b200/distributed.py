"""Multi-GPU partitioning of the hot path — only where it shards naturally.

One process per GPU (torchrun), `torch.distributed` for the plumbing:
  * predict over N_test: independent test points -> contiguous shard per rank, NO data-path collective;
    an optional all_gather returns the full (mean, variance) on every rank.
  * hyper-parameter / restart batches: independent fits -> restart r runs on rank r mod world; one
    all_gather of the scalar objectives (and gradients) at the end.
  * ONE large fit over all the GPUs (`ShardedLaplaceGP`): rows of K sharded for the Newton / CG iterations,
    block-column-cyclic Cholesky with NCCL panel broadcasts, test points sharded.  All of that runs inside the
    C ABI (pb_dist_laplace_fit / pb_dist_predict, csrc/dist.cu): the collectives are enqueued on CUDA streams from
    C++, Python only bootstraps the communicator (128-byte NCCL id through torch.distributed).
The reference has no distributed code at all; the first two helpers are the host logic of
BASELINE configs[4] and are backend-agnostic (nccl on GPUs; the CPU tests drive them with gloo and an injected
compute function).
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of range(n) for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def predict_sharded(predict_fn, X_test, gather=False, group=None):
    """Run `predict_fn(X_shard) -> (mean, var)` on this rank's shard of X_test.

    Returns (mean, var, (lo, hi)); with gather=True every rank gets the full vectors instead.
    """
    rank, world = _world(group)
    n = X_test.shape[0]
    lo, hi = shard_range(n, rank, world)
    mean, var = predict_fn(X_test[lo:hi])
    if not gather or world == 1:
        return mean, var, (lo, hi)
    sizes = [shard_range(n, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    pad = torch.zeros((2, width), dtype=mean.dtype, device=mean.device)
    pad[0, : hi - lo] = mean
    pad[1, : hi - lo] = var
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    full_m = torch.cat([out[r][0, : h - l] for r, (l, h) in enumerate(sizes)])
    full_v = torch.cat([out[r][1, : h - l] for r, (l, h) in enumerate(sizes)])
    return full_m, full_v, (0, n)


def restart_batch(evaluate_fn, parameter_list, group=None, device=None):
    """Evaluate `evaluate_fn(parameters) -> float | sequence of floats` for every entry of
    parameter_list, restart r on rank r mod world; returns a (len(parameter_list), k) tensor on every rank."""
    rank, world = _world(group)
    mine = list(range(rank, len(parameter_list), world))
    rows = []
    for r in mine:
        v = evaluate_fn(parameter_list[r])
        rows.append(torch.as_tensor(v, dtype=torch.float64).reshape(-1))
    k = rows[0].numel() if rows else 1
    per_rank = (len(parameter_list) + world - 1) // world
    local = torch.full((per_rank, k), float("nan"), dtype=torch.float64, device=device)
    for i, row in enumerate(rows):
        local[i] = row.to(local.device)
    if world == 1:
        return local[: len(parameter_list)]
    ks = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(ks, torch.tensor([k], dtype=torch.int64, device=local.device), group=group)
    k = int(max(int(t.item()) for t in ks))
    if local.shape[1] != k:
        local = torch.full((per_rank, k), float("nan"), dtype=torch.float64, device=device)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local, group=group)
    full = torch.empty((len(parameter_list), k), dtype=torch.float64, device=local.device)
    for r in range(len(parameter_list)):
        full[r] = out[r % world][r // world]
    return full


# ------------------------------------------------------------------------------------------------
# ONE fit across the GPUs (SURVEY.md §8e "large-N Cholesky"; BASELINE configs[3] at 2/4/8 GPUs)
# ------------------------------------------------------------------------------------------------
def exchange_unique_id(make_id, group=None, device=None):
    """Rank 0 calls `make_id() -> 128 bytes`; every rank returns the same bytes (one broadcast of a uint8 tensor
    through torch.distributed: NCCL needs a CUDA tensor, gloo a CPU one — `device` says which)."""
    rank, world = _world(group)
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    if world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class Communicator:
    """pb_comm: the library's own NCCL communicator (plus its communication / look-ahead streams) on the current
    CUDA device.  world == 1 needs no NCCL at all."""

    def __init__(self, group=None):
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = _world(group)
        self.group = group
        handle = C.c_void_p(0)
        if self.world > 1:
            def make_id():
                raw = (C.c_char * 128)()
                _lib.check(self.lib.pb_comm_unique_id(raw))
                return bytes(raw)
            uid = exchange_unique_id(make_id, group, "cuda")
            _lib.check(self.lib.pb_comm_create(uid, self.rank, self.world, C.byref(handle)))
        else:
            _lib.check(self.lib.pb_comm_create(None, 0, 1, C.byref(handle)))
        self.handle = handle

    def close(self):
        if self.handle:
            self.lib.pb_comm_destroy(self.handle)
            self.handle = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _sharded_class():
    from . import _lib, approximators as _appr
    from .linalg import _dev, _mat2, _ptr, _stream

    class ShardedLaplaceGP(_appr.LaplaceGP):
        """LaplaceGP (probit/approximators.py:213-277) for ONE training set partitioned over the ranks of a process
        group.  Same constructor and methods; every rank passes the SAME data and parameters and receives the same
        (weight, precision).  `predict` takes this rank's LOCAL test points (shard them with `shard_range`;
        `predict_global` does that and all-gathers the result).  Memory per GPU is N^2/G for the Gram rows plus
        N^2/G for the rank's block columns of the factor, so N beyond one GPU's 180 GB fits."""

        def __init__(self, data, prior, log_likelihood, grad_log_likelihood=None, hessian_log_likelihood=None,
                     comm=None, group=None, **kwargs):
            super().__init__(data, prior, log_likelihood, grad_log_likelihood, hessian_log_likelihood, **kwargs)
            self.comm = comm if comm is not None else Communicator(group)
            self.group = self.comm.group
            self.rank, self.world = self.comm.rank, self.comm.world

        def __repr__(self):
            return f"ShardedLaplaceGP(rank {self.rank} of {self.world})"

        def _workspace(self):
            if self._ws is None:
                self._ws_bytes = self.lib.pb_dist_workspace_bytes(self.N, self.D, self.world, self.rank, C.byref(self.options))
                self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device="cuda")
            return self._ws

        def _fit(self, parameters, final_factor):
            prob, keep = self._problem(parameters)
            ws = self._workspace()
            w = torch.empty(self.N, dtype=torch.float64, device="cuda")
            p = torch.empty_like(w)
            f = torch.empty_like(w)
            res = _lib.FitResult()
            self._gram_key, self._factor_key = None, None
            status = self.lib.pb_dist_laplace_fit(_stream(), self.comm.handle, C.byref(prob), float(self.tolerance),
                                                  int(self.maxiter), _ptr(ws), self._ws_bytes, _ptr(w), _ptr(p), _ptr(f),
                                                  C.byref(res), C.byref(self.options))
            self.last_result = res
            _lib.check(status)
            if final_factor:
                self.last_logdet = self._factor_predict(prob, p, w, None, float(self.jitter), want_logdet=True)[2]
            del keep
            return w, p, f

        def _chunk_rows(self, n_test):
            if self.predict_chunk:
                return max(1, min(int(self.predict_chunk), max(n_test, 1)))
            free, _ = torch.cuda.mem_get_info()
            budget = min(int(free * 0.7), 64 << 30)
            ld = (self.N + 15) // 16 * 16
            return max(1, min(max(n_test, 1), budget // (8 * ld + 16 * self.D + 512)))

        def _factor_predict(self, prob, precision, weight, X_test, jitter, want_logdet=False, variance=True):
            ws = self._workspace()
            n_test = 0 if X_test is None else X_test.shape[0]
            key = (self._spec_key(prob.kernel), float(jitter))
            reuse = (self._factor_key is not None and self._factor_key[0] == key
                     and torch.equal(self._factor_key[1], precision) and not want_logdet)
            chunk = self._chunk_rows(n_test)
            sbytes = self.lib.pb_dist_predict_scratch_bytes(self.N, self.D, min(chunk, max(n_test, 1))) if n_test else 0
            scratch = torch.empty(max(sbytes, 256), dtype=torch.uint8, device="cuda")
            mean = torch.empty(n_test, dtype=torch.float64, device="cuda")
            var = torch.empty(n_test, dtype=torch.float64, device="cuda") if variance else None
            logdet, info = C.c_double(float("nan")), C.c_int32(0)
            self._factor_key = None
            _lib.check(self.lib.pb_dist_predict(
                _stream(), self.comm.handle, C.byref(prob), _ptr(precision), _ptr(weight), _ptr(ws), self._ws_bytes,
                float(jitter), int(reuse), _ptr(X_test) if n_test else None, n_test, chunk, _ptr(scratch), sbytes,
                _ptr(mean) if n_test else None, _ptr(var) if (variance and n_test) else None,
                C.byref(logdet) if want_logdet else None, C.byref(info), C.byref(self.options)))
            if variance or want_logdet:
                self._factor_key = (key, precision.clone())
            return mean, var, logdet.value

        def predict(self, X_test, parameters, weight, precision, variance=True):
            """(mean, variance) at this rank's LOCAL test points (approximators.py:154-180).  Every rank must call
            it (the panels of the factor are broadcast to all ranks), possibly with zero rows."""
            prob, keep = self._problem(parameters)
            weight = _dev(weight).reshape(-1)
            precision = _dev(precision).reshape(-1)
            X_test = _mat2(X_test) if X_test is not None and len(X_test) else None
            if X_test is not None and X_test.shape[1] != self.D:
                raise ValueError("X_test has the wrong input dimension")
            mean, var, _ = self._factor_predict(prob, precision, weight, X_test, 0.0, variance=variance)
            del keep
            return mean.to(self.out_dtype), (var.to(self.out_dtype) if var is not None else None)

        def predict_global(self, X_test, parameters, weight, precision):
            """Full-length (mean, variance) on every rank: shard X_test, predict locally, all-gather."""
            def fn(Xs):
                return self.predict(Xs, parameters, weight, precision)
            m, v, _ = predict_sharded(fn, _mat2(X_test), gather=True, group=self.group)
            return m, v

        def posterior_mean(self, weight, parameters):
            """K @ weight without a resident K: the fused on-the-fly matvec over this rank's training rows + all-gather."""
            prob, keep = self._problem(parameters)
            weight = _dev(weight).reshape(-1)

            def fn(Xs):
                m = self._factor_predict(prob, torch.ones_like(weight), weight, Xs, 0.0, variance=False)[0]
                return m, torch.zeros_like(m)
            m, _, _ = predict_sharded(fn, self.X, gather=True, group=self.group)
            del keep
            return m

        def objective(self):
            """objective_LA (Laplace.py:12-30) with the log-determinant from the block-cyclic factor."""
            def obj(parameters):
                self._fit(parameters, final_factor=True)
                r = self.last_result
                return -r.sum_ll + 0.5 * r.ftw + self.last_logdet
            return obj

        def value_and_grad(self):
            raise NotImplementedError("ShardedLaplaceGP: gradients need the full inverse; run hyper-parameter batches "
                                      "one per GPU with restart_batch instead (SURVEY.md §8e)")

        def predict_covariance(self, *a, **k):
            raise NotImplementedError("ShardedLaplaceGP: predict_covariance is single-GPU only")

    return ShardedLaplaceGP


def __getattr__(name):
    if name == "ShardedLaplaceGP":
        cls = _sharded_class()
        globals()["ShardedLaplaceGP"] = cls
        return cls
    raise AttributeError(name)
