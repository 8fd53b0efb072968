#!/bin/bash
# Round-2 single-GPU validation + measurements (one gpurun call).  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/a_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench_1gpu.json 2> gpurun_out/a_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python tools/ncu_hbm_kernels.py --time > gpurun_out/a_hbm_timed.json 2> gpurun_out/a_hbm_timed.err; echo "hbm timed rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'likelihood_kernel|gram_sym_kernel|gram_matvec_kernel|predictive_kernel' -c 6 -f -o gpurun_out/a_hbm \
  python tools/ncu_hbm_kernels.py > gpurun_out/a_hbm_ncu.log 2>&1; echo "hbm ncu rc=$?"
tail -3 gpurun_out/a_pytest.log; tail -2 gpurun_out/a_smoke.log; head -c 600 gpurun_out/a_bench_1gpu.json
timeout 600 python tools/ozaki_bench.py 16384 32768 > gpurun_out/a_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/a_ozaki.log
