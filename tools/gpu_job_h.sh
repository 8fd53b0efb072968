#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 16384 32768 --test-n 2048 --json gpurun_out/h_dist2.json > gpurun_out/h_dist2.log 2>&1; echo "dist2 rc=$?"; grep "rank 0" gpurun_out/h_dist2.log | cut -c1-1000; tail -3 gpurun_out/h_dist2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/h_bench_2gpu.json 2> gpurun_out/h_bench_2gpu.err; echo "bench2 rc=$?"; head -c 300 gpurun_out/h_bench_2gpu.json; tail -3 gpurun_out/h_bench_2gpu.err | cut -c1-300
