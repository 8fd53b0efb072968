"""Communication chain of the block-cyclic Cholesky in isolation (torchrun, one rank per GPU): the nblk shrinking panel
broadcasts of one N x N factorisation issued back to back through NCCL, no compute.  Tells how much of the multi-GPU
factorisation time is the panels' critical path.   usage: bcast_bench.py [N] [NB]"""
import os, sys, time
import torch, torch.distributed as dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
buf = torch.empty(n * nb + 64 * nb, dtype=torch.float64, device="cuda")
nblk = (n + nb - 1) // nb
def chain():
    for k in range(nblk):
        cnt = nb * 64 + (n - k * nb) * nb
        dist.broadcast(buf[:cnt], src=k % world)
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    chain()
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    if rank == 0:
        total = sum(nb * 64 + (n - k * nb) * nb for k in range(nblk)) * 8
        print(f"world {world} N {n} NB {nb}: {nblk} panel broadcasts {total / 1e9:.1f} GB in {t * 1e3:.1f} ms = {total / t / 1e9:.0f} GB/s", flush=True)
# one full-size panel, repeated
cnt = nb * 64 + n * nb
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        dist.broadcast(buf[:cnt], src=0)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
    if rank == 0:
        print(f"  one {cnt * 8 / 1e6:.0f} MB broadcast: {t * 1e3:.3f} ms = {cnt * 8 / t / 1e9:.0f} GB/s", flush=True)
dist.destroy_process_group()
