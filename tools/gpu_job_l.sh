#!/bin/bash
mkdir -p gpurun_out
PB_OZ_TILE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki or potrf or int8" > gpurun_out/l_pytest.log 2>&1; rc=$?; echo "pytest(tile128) rc=$rc"; tail -12 gpurun_out/l_pytest.log | cut -c1-400
PB_OZ_TILE=1 timeout 600 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/l_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/l_ozaki.log
