"""CPU baseline scan: the oracle port on the host cores at several N, both in the reference's
literal operation sequence (dense Jacobian + LU) and in the Cholesky form the GPU path uses, so that the
algorithmic and the hardware part of the speed-up can be separated.  Fits t = a N^3 and extrapolates to 65536."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import cpu_problem, NOISE_VARIANCE
from oracle import approximators as OA, kernels as OK, utilities as OU

out = {"cores": os.cpu_count(), "rows": []}
try:
    from threadpoolctl import threadpool_info
    out["blas"] = [{k: i.get(k) for k in ("internal_api", "version", "num_threads")} for i in threadpool_info()]
except Exception:
    pass
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
for n in sizes:
    X, y, cut, Xs = cpu_problem(n, min(n, 1024))
    params = (1.0, (float(np.sqrt(NOISE_VARIANCE)), cut))
    row = {"N": n, "N_test": len(Xs)}
    for form in ("lu_jacobian", "cholesky_B"):
        gp = OA.LaplaceGP((X, y), lambda l: 1.0 * OK.Matern12().stretch(l), OU.log_probit_likelihood, newton_form=form)
        t0 = time.perf_counter()
        w, p = gp.approximate_posterior(params)
        t1 = time.perf_counter()
        gp.predict(Xs, params, w, p)
        t2 = time.perf_counter()
        row[form] = {"fit_s": t1 - t0, "predict_s": t2 - t1, "iterations": len(gp.trace)}
    out["rows"].append(row)
    print(row, flush=True)
for form in ("lu_jacobian", "cholesky_B"):
    a = np.mean([r[form]["fit_s"] / r["N"] ** 3 for r in out["rows"][-2:]])
    out[f"{form}_fit_extrapolated_to_65536_s"] = float(a * 65536.0 ** 3)
print(json.dumps(out, indent=1))
