#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ozaki or potrf or int8" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/d_pytest.log
timeout 900 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/d_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -12 gpurun_out/d_ozaki.log
