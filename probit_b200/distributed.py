"""Multi-GPU partitioning of the hot path — only where it shards naturally.

One process per GPU (torchrun), `torch.distributed` for the plumbing:
  * predict over N_test: independent test points -> contiguous shard per rank, NO data-path collective;
    an optional all_gather returns the full (mean, variance) on every rank.
  * hyper-parameter / restart batches: independent fits -> restart r runs on rank r mod world; one
    all_gather of the scalar objectives (and gradients) at the end.
The reference has no distributed code at all; these helpers are the host logic of
BASELINE configs[4].  They are backend-agnostic (nccl on GPUs; the CPU tests drive them with gloo
and an injected compute function).
"""
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
# Large-N Cholesky across GPUs: 1-D block-column-cyclic, right-looking, one-panel
# look-ahead, panels broadcast over the process group (NCCL on NVLink/NVSwitch).
#
# Why 1-D and not a 2-D grid: on NVSwitch every GPU receives every panel at full link bandwidth, and
# the whole factorisation moves only 8*N^2/2 bytes per GPU (17 GB at N=65536, ~25 ms at 700 GB/s)
# against N^3/(3G) flops of trailing update (>= 0.34 s at G=8), so the broadcast volume that a 2-D
# layout would save is already hidden under the update; the 1-D layout keeps every trailing update a
# single large DMMA GEMM per owned block column.  Because every rank sees every panel, each rank also
# assembles the complete factor for free (no gather), after which triangular solves and predict run
# replicated / sharded over test points.
# ------------------------------------------------------------------------------------------------
class BlockCyclicCholesky:
    def __init__(self, n, ops, nb=512, group=None):
        self.n, self.nb, self.ops, self.group = int(n), int(nb), ops, group
        self.rank, self.world = _world(group)
        self.nblk = (self.n + self.nb - 1) // self.nb
        self.owned = [j for j in range(self.nblk) if j % self.world == self.rank]
        self.ld_loc = max(len(self.owned), 1) * self.nb
        self.Aloc = ops.empty(self.n, self.ld_loc)                 # owned block columns, side by side
        self.P = [ops.empty(self.n, self.nb), ops.empty(self.n, self.nb)]   # double-buffered panel (contiguous)

    def width(self, j):
        return min(self.nb, self.n - j * self.nb)

    def owner(self, j):
        return j % self.world

    def local_block(self, j):
        """View of block column j (rows j0.., its own width) inside the local storage."""
        jl = j // self.world
        j0 = j * self.nb
        return self.Aloc[j0:, jl * self.nb: jl * self.nb + self.width(j)]

    def panel_view(self, k):
        """(n - k0, nb) contiguous panel buffer; only the first width(k) columns are meaningful."""
        m = self.n - k * self.nb
        return self.P[k % 2].reshape(-1)[: m * self.nb].view(m, self.nb)

    def _bcast(self, k):
        buf = self.panel_view(k)
        if self.world == 1:
            return None
        return dist.broadcast(buf, src=dist.get_global_rank(self.group, self.owner(k)) if self.group is not None
                              else self.owner(k), group=self.group, async_op=True)

    def _produce(self, k):
        """Owner only: factor panel k in place (diagonal block + rows below) and pack it for the broadcast."""
        w = self.width(k)
        blk = self.local_block(k)
        info = self.ops.potrf_panel(blk, w)
        self.panel_view(k)[:, :w].copy_(blk)
        return info

    def _update(self, j, k, Pk):
        """Block column j -= P_k[rows >= j0] * P_k[rows of block j]^T."""
        off = (j - k) * self.nb
        wj, wk = self.width(j), self.width(k)
        self.ops.gemm_nt(Pk[off:, :wk], Pk[off: off + wj, :wk], self.local_block(j), alpha=-1.0, beta=1.0)

    def factor(self, fill_block, write_panel):
        """fill_block(j0, w, out): write rows j0.. of columns [j0, j0+w) of the SPD matrix into `out`.
        write_panel(k0, w, panel): called on EVERY rank for EVERY factored panel ((n-k0) x w, rows k0..).
        Returns the list of (k0, info) pairs produced by this rank's panel factorisations."""
        for j in self.owned:
            fill_block(j * self.nb, self.width(j), self.local_block(j))
        infos = []
        if self.owner(0) == self.rank:
            infos.append((0, self._produce(0)))
        work = self._bcast(0)
        for k in range(self.nblk):
            if work is not None:
                work.wait()
            Pk = self.panel_view(k)
            write_panel(k * self.nb, self.width(k), Pk[:, : self.width(k)])
            nxt = k + 1
            if nxt < self.nblk:
                if self.owner(nxt) == self.rank:          # look-ahead: next panel first, then ship it
                    self._update(nxt, k, Pk)
                    infos.append((nxt * self.nb, self._produce(nxt)))
                work = self._bcast(nxt)                   # receivers post before their own updates
            for j in self.owned:
                if j > k and j != nxt:
                    self._update(j, k, Pk)
        return infos


class TorchCholeskyOps:
    """CUDA ops of BlockCyclicCholesky: the product's own C ABI (pb_potrf, pb_trsm_right_lt, pb_gemm_nt)."""

    def __init__(self):
        from . import _lib, linalg
        self.lib, self.linalg = _lib.load(), linalg
        self._ws = None

    def empty(self, rows, cols):
        return torch.empty((rows, cols), dtype=torch.float64, device="cuda")

    def potrf_panel(self, blk, w):
        import ctypes as C
        lib, la = self.lib, self.linalg
        need = lib.pb_potrf_workspace_bytes(w)
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.empty(need // 8, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = la._stream()
        ld = blk.stride(0)
        from ._lib import check
        check(lib.pb_potrf(st, la._ptr(blk), w, ld, la._ptr(self._ws), need, la._ptr(info)))
        m = blk.shape[0] - w
        if m > 0:
            below = blk[w:, :]
            check(lib.pb_trsm_right_lt(st, la._ptr(blk), w, ld, la._ptr(self._ws), la._ptr(below), m, ld))
        return info

    def gemm_nt(self, A, B, C_out, alpha, beta):
        self.linalg.gemm_nt(A, B, C_out, alpha=alpha, beta=beta, lower_only=False)


def row_shard(n, rank, world):
    """Equal-sized (padded) row chunks for an all-gather: returns (lo, hi, chunk); ranks past the end get lo == hi."""
    chunk = -(-n // world)
    return min(n, rank * chunk), min(n, (rank + 1) * chunk), chunk


def sharded_matvec(local_product, n, group=None, device=None, buffers=None):
    """y = K x with the rows of K split over the ranks.  `local_product(lo, hi, out)` writes K[lo:hi] @ x into
    out[:hi - lo]; one all_gather_into_tensor of 8 * chunk bytes per rank returns the same y (length n) everywhere.
    `buffers` (dict) caches the padded send / receive tensors between calls."""
    rank, world = _world(group)
    lo, hi, chunk = row_shard(n, rank, world)
    buffers = buffers if buffers is not None else {}
    if buffers.get("chunk") != (chunk, world):
        buffers["chunk"] = (chunk, world)
        buffers["loc"] = torch.zeros(chunk, dtype=torch.float64, device=device)
        buffers["full"] = torch.zeros(world * chunk, dtype=torch.float64, device=device)
    loc, full = buffers["loc"], buffers["full"]
    if hi > lo:
        local_product(lo, hi, loc)
    if world > 1:
        dist.all_gather_into_tensor(full, loc, group=group)
        return full[:n]
    return loc[:n]


class DistributedFactorization:
    """Installs the block-cyclic Cholesky as the factorisation of an approximator's fit / predict drivers
    (pb_set_factor_callback).  Every rank must run the same fit on the same data (replicas)."""

    def __init__(self, approximator, group=None, nb=512, shard_matvec=None):
        """shard_matvec: also replace y = K x of the Newton / CG iterations by a row-sharded product followed by an
        all-gather (pb_set_matvec_callback); default: whenever the group has more than one rank."""
        import ctypes as C
        from . import _lib, linalg
        self.gp, self.group, self.lib, self.la = approximator, group, _lib.load(), linalg
        self.chol = BlockCyclicCholesky(approximator.N, TorchCholeskyOps(), nb=nb, group=group)
        self.calls = 0
        self.matvec_calls = 0
        self.shard_matvec = (self.chol.world > 1) if shard_matvec is None else bool(shard_matvec)
        self._mv_buf = None

        def matvec(user, stream, K, n, ldk, x, y):
            try:
                self._matvec(K, n, ldk, x, y)
                return _lib.PB_OK
            except Exception as exc:
                self.error = exc
                return _lib.PB_ERR_CUDA

        self._mv_cb = _lib.MATVEC_FN(matvec)

        def callback(user, stream, K, n, ldk, s, a, jitter, L, ldl, pws, pws_bytes, info_dev):
            try:
                self._factor(K, n, ldk, s, a, jitter, L, ldl, pws, pws_bytes, info_dev)
                return _lib.PB_OK
            except Exception as exc:          # never let an exception cross the C boundary
                self.error = exc
                return _lib.PB_ERR_CUDA

        self._cb = _lib.FACTOR_FN(callback)
        self.error = None

    def __enter__(self):
        import ctypes as C
        self.lib.pb_set_factor_callback(C.cast(self._cb, C.c_void_p), None)
        if self.shard_matvec:
            self.lib.pb_set_matvec_callback(C.cast(self._mv_cb, C.c_void_p), None)
        return self

    def __exit__(self, *exc):
        self.lib.pb_set_factor_callback(None, None)
        self.lib.pb_set_matvec_callback(None, None)
        return False

    def _matvec(self, K, n, ldk, x, y):
        """y = K x with the rows of K split evenly over the ranks: each rank streams n/G rows (the product is HBM
        bound, so this is the 1/G of the time) and one all-gather of 8n bytes puts the same y on every rank."""
        import ctypes as C
        lib, la = self.lib, self.la

        def local_product(lo, hi, out):
            _check(lib.pb_gemv(la._stream(), C.c_void_p(K + lo * ldk * 8), hi - lo, n, ldk, C.c_void_p(x), la._ptr(out)))

        if self._mv_buf is None:
            self._mv_buf = {}
        full = sharded_matvec(local_product, n, self.group, "cuda", self._mv_buf)
        ws = self.gp._workspace()
        off = y - ws.data_ptr()
        ws[off: off + n * 8].view(torch.float64).copy_(full)
        self.matvec_calls += 1

    def _view(self, ptr, rows, ld):
        ws = self.gp._workspace()
        off = ptr - ws.data_ptr()
        return ws[off: off + rows * ld * 8].view(torch.float64).view(rows, ld)

    def _factor(self, K, n, ldk, s, a, jitter, L, ldl, pws, pws_bytes, info_dev):
        import ctypes as C
        lib, la = self.lib, self.la
        st = la._stream()
        Lv = self._view(L, n, ldl)[:, :n]

        def fill_block(j0, w, out):
            _check(lib.pb_transform_block(st, C.c_void_p(K), ldk, C.c_void_p(s) if s else None, a, jitter, j0, j0,
                                          n - j0, w, la._ptr(out), out.stride(0)))

        def write_panel(k0, w, panel):
            Lv[k0:, k0: k0 + w].copy_(panel)

        infos = self.chol.factor(fill_block, write_panel)
        glob = torch.zeros(1, dtype=torch.int32, device="cuda")
        for k0, info in infos:
            glob = torch.where((glob == 0) & (info > 0), info + k0, glob)
        if self.chol.world > 1:
            big = torch.where(glob == 0, torch.full_like(glob, 2**31 - 1), glob)
            dist.all_reduce(big, op=dist.ReduceOp.MIN, group=self.group)
            glob = torch.where(big == 2**31 - 1, torch.zeros_like(big), big)
        ws = self.gp._workspace()
        off = info_dev - ws.data_ptr()
        ws[off: off + 4].view(torch.int32).copy_(glob)
        _check(lib.pb_rebuild_solve_workspace(st, C.c_void_p(L), n, ldl, C.c_void_p(pws), pws_bytes))
        self.calls += 1


def _check(status):
    from . import _lib
    _lib.check(status)
