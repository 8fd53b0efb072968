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
