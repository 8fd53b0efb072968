"""Newton-phase policy sweep at the north-star size: fit-only time and CG iterations per Nystrom rank.
Usage: python tools/newton_sweep.py [N] [rank ...]   (rank 0 = stale-factor policy)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from probit_b200 import _lib, approximators as PA, kernels as PK, utilities as PU
from probit_b200.datasets import device_latent_sampler, generate_ordinal_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ranks = [int(a) for a in sys.argv[2:]] or [0, 1024, 2048, 4096, 8192]
X, g, y, cut = generate_ordinal_data(1, n, 4, 5, 0.4, device_latent_sampler(1.0 * PK.Matern12().stretch(1.0), 1e-6))
torch.cuda.empty_cache()
out = {}
for fam, ell in (("matern12", 1.0), ("matern12", 0.25), ("matern12", 4.0), ("eq", 1.0), ("eq", 0.3)):
    base = PK.Matern12 if fam == "matern12" else PK.EQ
    gp = PA.LaplaceGP((torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()), lambda l: 1.0 * base().stretch(l),
                      PU.log_probit_likelihood, tolerance=1e-5)
    params = (ell, (float(np.sqrt(0.4)), cut))
    for r in ranks:
        gp.options.laplace_nystrom_rank = r
        best = 1e30
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); w, p = gp.approximate_posterior(params); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res = gp.last_result
        out["%s_l%g_r%d" % (fam, ell, r)] = dict(fit_ms=best, newton=res.iterations, potrf=res.factorizations, cg=res.pcg_iterations)
        print(fam, ell, "rank", r, "fit ms %.1f" % best, "newton", res.iterations, "potrf", res.factorizations, "cg", res.pcg_iterations, flush=True)
    del gp
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "newton_sweep.json"), "w"), indent=1)
