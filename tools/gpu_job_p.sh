#!/bin/bash
mkdir -p gpurun_out
export PB_OZ_CLUSTER=1
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki" > gpurun_out/p_pytest.log 2>&1; rc=$?; echo "pytest(cluster, ozaki) rc=$rc"; tail -8 gpurun_out/p_pytest.log | cut -c1-400
if [ $rc -ne 0 ]; then exit 0; fi
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "potrf or int8" > gpurun_out/p_pytest2.log 2>&1; rc=$?; echo "pytest(cluster, potrf) rc=$rc"; tail -5 gpurun_out/p_pytest2.log | cut -c1-400
timeout 400 python tools/ozaki_bench.py 16384 32768 65536 > gpurun_out/p_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -8 gpurun_out/p_ozaki.log
for cfg in "16384 512" "16384 1024"; do PB_OZ_TIMING=5000 timeout 120 python tools/oz_timeline.py $cfg 2>&1 | tail -1; done | tee gpurun_out/p_timeline.log
