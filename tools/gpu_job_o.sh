#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki or potrf or int8" > gpurun_out/o_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/o_pytest.log | cut -c1-400
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/o_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/o_ozaki.log
for cfg in "16384 512" "16384 1024"; do PB_OZ_TIMING=5000 timeout 120 python tools/oz_timeline.py $cfg 2>&1 | tail -1; done | tee gpurun_out/o_timeline.log
