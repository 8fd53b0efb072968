#!/bin/bash
# 8-GPU: strong-scaled fit bench, predict over 10^6 points, one traced factorisation
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/j_bench_8gpu.json 2> gpurun_out/j_bench_8gpu.err; echo "bench8 rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/j_bench_8gpu.json | head -2; grep -o '"stages": {[^}]*}' gpurun_out/j_bench_8gpu.json
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --workload predict --test-n 1000000 --steps 1 --warmup 1 > gpurun_out/j_predict_8gpu.json 2> gpurun_out/j_predict_8gpu.err; echo "predict8 rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/j_predict_8gpu.json | head -1
PB_DIST_TRACE=gpurun_out/j_trace timeout 300 $TR --master-port 29523 tools/dist_check.py 65536 --single-max 0 --reps 2 --test-n 4096 --json gpurun_out/j_dist8.json > gpurun_out/j_dist8.log 2>&1; echo "trace rc=$?"; grep "rank 0" gpurun_out/j_dist8.log | cut -c1-400
