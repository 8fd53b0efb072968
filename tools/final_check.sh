#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/z_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/z_smoke.log | cut -c1-300
