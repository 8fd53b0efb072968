#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki" > gpurun_out/q_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/q_pytest.log | cut -c1-300
for cfg in "16384 512" "16384 1024"; do PB_OZ_TIMING=5000 timeout 120 python tools/oz_timeline.py $cfg 2>&1 | tail -1; done | tee gpurun_out/q_timeline.log
timeout 300 python tools/ozaki_bench.py 16384 > gpurun_out/q_ozaki.log 2>&1; tail -6 gpurun_out/q_ozaki.log
