#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/f_pytest.log
timeout 900 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/f_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/f_ozaki.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo "bench rc=$?"; head -c 300 gpurun_out/f_bench_1gpu.json; tail -3 gpurun_out/f_bench_1gpu.err
