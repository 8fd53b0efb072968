#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/k_pytest.log
timeout 300 python tools/ncu_hbm_kernels.py --time > gpurun_out/k_hbm_timed.json 2> gpurun_out/k_hbm_timed.err; echo "hbm timed rc=$?"; cat gpurun_out/k_hbm_timed.json
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'likelihood_kernel|gram_sym_kernel|predictive_kernel|gram_matvec_kernel' -c 5 -f -o gpurun_out/k_hbm \
  python tools/ncu_hbm_kernels.py > gpurun_out/k_hbm_ncu.log 2>&1; echo "hbm ncu rc=$?"
timeout 300 python tools/ozaki_bench.py 16384 > gpurun_out/k_ozaki.log 2>&1; tail -3 gpurun_out/k_ozaki.log
