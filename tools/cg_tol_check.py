"""Accuracy of the CG Newton policy against factor-every-step as a function of the CG tolerance (one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from probit_b200 import _lib, approximators as PA, kernels as PK, utilities as PU
from probit_b200.datasets import device_latent_sampler, generate_ordinal_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
X, g, y, cut = generate_ordinal_data(1, n, 4, 5, 0.4, device_latent_sampler(1.0 * PK.Matern12().stretch(1.0), 1e-6))
torch.cuda.empty_cache()
gp = PA.LaplaceGP((torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()), lambda l: 1.0 * PK.Matern12().stretch(l),
                  PU.log_probit_likelihood, tolerance=1e-5)
params = (1.0, (float(np.sqrt(0.4)), cut))
gp.options.laplace_pcg_min_n = 1 << 40
w0, p0 = gp.approximate_posterior(params)
print("factor every step: newton", gp.last_result.iterations, "error", gp.last_result.error, flush=True)
gp.options.laplace_pcg_min_n = 0
for tol in (1e-1, 1e-2, 1e-3, 1e-4):
    gp.options.laplace_cg_tol = tol
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w, p = gp.approximate_posterior(params); e1.record(); e1.synchronize()
    r = gp.last_result
    print("tol %.0e: fit %.0f ms, newton %d, potrf %d, cg %d, weight rel diff %.2e, precision rel diff %.2e, final error %.3e" % (
        tol, e0.elapsed_time(e1), r.iterations, r.factorizations, r.pcg_iterations,
        ((w - w0).norm() / w0.norm()).item(), ((p - p0).norm() / p0.norm()).item(), r.error), flush=True)
