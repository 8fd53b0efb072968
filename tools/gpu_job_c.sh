#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c_bench_1gpu.json 2> gpurun_out/c_bench_1gpu.err; echo "bench rc=$?"; head -c 400 gpurun_out/c_bench_1gpu.json; tail -3 gpurun_out/c_bench_1gpu.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'likelihood_kernel|gram_sym_kernel|predictive_kernel' -c 4 -f -o gpurun_out/c_hbm \
  python tools/ncu_hbm_kernels.py > gpurun_out/c_hbm_ncu.log 2>&1; echo "hbm ncu rc=$?"
