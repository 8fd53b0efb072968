#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki or potrf or int8" > gpurun_out/i_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/i_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/i_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/i_ozaki.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_1gpu.json 2> gpurun_out/i_bench_1gpu.err; echo "bench rc=$?"; head -c 250 gpurun_out/i_bench_1gpu.json; tail -3 gpurun_out/i_bench_1gpu.err
