#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 1 -c 1 -f -o gpurun_out/e_oz python tools/ozaki_once.py 8192 1024 > gpurun_out/e_oz_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/e_oz_ncu.log
