#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dist_check.py 16384 --test-n 2048 --json gpurun_out/g_dist1.json > gpurun_out/g_dist1.log 2>&1; echo "dist world1 rc=$?"; tail -5 gpurun_out/g_dist1.log | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -k "sharded" > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g_pytest.log
