"""Executable model (test infrastructure) of the multi-GPU schedule in probit_b200/csrc/dist.cu.

The product runs this schedule in C++ on CUDA streams with NCCL; it cannot execute without GPUs.  This file restates
the SAME index arithmetic — block-column-cyclic ownership, the storage of an owned block column (`col`), the panel
message (leaf part omitted), the look-ahead order (column k+1 on the owner first, then the owned columns >= k+2
nearest first), the three rotating panel buffers, and the per-panel hook that carries the test rows V along
(V_k <- V_k L_kk^-T, V[:, k1:] -= V_k L[k1:, k]^T) — over torch CPU tensors and a torch.distributed group, so that
the world_size-2 gloo tests pin the algebra and the partitioning of what dist.cu does (bc_factor / bc_stream /
pb_dist_predict).  Nothing in probit_b200/ imports it.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class BlockCyclicModel:
    def __init__(self, n, nb, group=None):
        self.n, self.nb, self.group = int(n), int(nb), group
        self.me, self.world = _world(group)
        self.nblk = (self.n + self.nb - 1) // self.nb
        self.owned_max = (self.nblk + self.world - 1) // self.world
        self.ld_loc = self.owned_max * self.nb
        self.Aloc = torch.zeros(self.n, self.ld_loc, dtype=torch.float64)          # ws.B(): n x ld_loc
        self.panels = [torch.zeros(self.n, self.nb, dtype=torch.float64) for _ in range(3)]
        self.log = []                                                              # (event, k) in issue order

    def width(self, k):
        return min(self.nb, self.n - k * self.nb)

    def owner(self, k):
        return k % self.world

    def col(self, j):
        """Bc::col: rows j0.., columns of block column j inside the local storage."""
        j0 = j * self.nb
        jl = (j // self.world) * self.nb
        return self.Aloc[j0:, jl: jl + self.width(j)]

    def panel(self, k):
        return self.panels[k % 3][: self.n - k * self.nb]

    # -- pieces of dist.cu ---------------------------------------------------------------------------------------
    def fill(self, entry):
        for j in range(self.me, self.nblk, self.world):
            j0 = j * self.nb
            rows = torch.arange(j0, self.n)
            cols = torch.arange(j0, j0 + self.width(j))
            self.col(j).copy_(entry(rows[:, None], cols[None, :]))

    def update_column(self, j, k):
        P = self.panel(k)
        off = (j - k) * self.nb
        wj, wk = self.width(j), self.width(k)
        self.col(j).sub_(P[off:, :wk] @ P[off: off + wj, :wk].T)

    def factor_column(self, k):
        w = self.width(k)
        A = self.col(k)
        d = torch.tril(A[:w, :w])
        L = torch.linalg.cholesky(d + torch.tril(d, -1).T)
        A[:w, :w] = L
        if A.shape[0] > w:
            A[w:, :] = torch.linalg.solve_triangular(L, A[w:, :].T, upper=False).T
        self.panel(k)[:, :w].copy_(A)                                             # pack_panel
        self.log.append(("factor", k))

    def ship(self, k):
        if self.world > 1:
            src = self.owner(k) if self.group is None else dist.get_global_rank(self.group, self.owner(k))
            dist.broadcast(self.panels[k % 3], src=src, group=self.group)
        self.log.append(("recv", k))

    def factor(self, entry, hook=None):
        """bc_factor: entry(i, j) gives the SPD matrix; hook(k, k0, w, P) sees every panel on every rank."""
        self.fill(entry)
        if self.owner(0) == self.me:
            self.factor_column(0)
        self.ship(0)
        for k in range(self.nblk):
            nx = k + 1
            if nx < self.nblk:
                if self.owner(nx) == self.me:
                    self.update_column(nx, k)
                    self.factor_column(nx)
                self.ship(nx)
            first = k + 2
            first += ((self.me - first) % self.world + self.world) % self.world
            for j in range(first, self.nblk, self.world):
                assert self.owner(j) == self.me and j >= k + 2
                self.update_column(j, k)
            if hook is not None:
                hook(k, k * self.nb, self.width(k), self.panel(k))

    def stream(self, hook):
        """bc_stream: re-broadcast the stored factor panel by panel."""
        for k in range(self.nblk):
            if self.owner(k) == self.me:
                self.panel(k)[:, : self.width(k)].copy_(self.col(k))
            self.ship(k)
            hook(k, k * self.nb, self.width(k), self.panel(k))


def make_apply_hook(V, n):
    """The V-rows hook of pb_dist_predict: after all panels V holds V L^-T."""
    def hook(k, k0, w, P):
        Lkk = torch.tril(P[:w, :w])
        V[:, k0:k0 + w] = torch.linalg.solve_triangular(Lkk, V[:, k0:k0 + w].T, upper=False).T
        k1 = k0 + w
        if k1 < n:
            V[:, k1:] -= V[:, k0:k0 + w] @ P[w:, :w].T
    return hook
