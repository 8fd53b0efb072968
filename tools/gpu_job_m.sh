#!/bin/bash
mkdir -p gpurun_out
PB_OZ_RED=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ozaki or int8" > gpurun_out/m_pytest.log 2>&1; rc=$?; echo "pytest(red) rc=$rc"; tail -5 gpurun_out/m_pytest.log | cut -c1-400
PB_OZ_RED=1 timeout 600 python tools/ozaki_bench.py 16384 65536 > gpurun_out/m_ozaki.log 2>&1; echo "ozaki rc=$?"; tail -7 gpurun_out/m_ozaki.log
