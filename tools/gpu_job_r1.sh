#!/bin/bash
# final single-GPU validation and records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r1_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1_smoke.log | cut -c1-300
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; echo "bench rc=$?"; head -c 200 gpurun_out/r1_bench_1gpu.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-comparators > gpurun_out/r1_ncu_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r1_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 1 -c 1 -f -o gpurun_out/r1_oz python tools/ozaki_once.py 8192 1024 > gpurun_out/r1_oz_ncu.log 2>&1; echo "ncu oz rc=$?"
