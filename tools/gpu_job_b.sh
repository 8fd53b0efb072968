#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/b_pytest.log
CUDA_LAUNCH_BLOCKING=1 timeout 600 python tools/bench_diag.py > gpurun_out/b_diag.log 2>&1; echo "diag rc=$?"; grep "diag\|rror" gpurun_out/b_diag.log | tail -25
timeout 300 python tools/ncu_hbm_kernels.py --time > gpurun_out/b_hbm_timed.json 2> gpurun_out/b_hbm_timed.err; echo "hbm timed rc=$?"; cat gpurun_out/b_hbm_timed.json
