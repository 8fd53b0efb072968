#!/bin/bash
# usage: gpu_job_multi.sh <ngpu> [fit] [predict] [restarts]
n=$1; shift
mkdir -p gpurun_out
if [ "$n" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n"; fi
for w in "$@"; do
  case $w in
    fit) timeout 600 $TR bench.py --gpus $n --steps 3 --warmup 2 --no-cpu-baseline --no-comparators > gpurun_out/s_bench_${n}gpu.json 2> gpurun_out/s_bench_${n}gpu.err; echo "fit$n rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/s_bench_${n}gpu.json | head -1; grep -o '"stages": {[^}]*}' gpurun_out/s_bench_${n}gpu.json;;
    predict) timeout 600 $TR bench.py --gpus $n --workload predict --test-n 1000000 --steps 1 --warmup 1 > gpurun_out/s_predict_${n}gpu.json 2> gpurun_out/s_predict_${n}gpu.err; echo "predict$n rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/s_predict_${n}gpu.json | head -1;;
    restarts) timeout 600 $TR bench.py --gpus $n --workload restarts --restarts 16 --steps 1 --warmup 0 > gpurun_out/s_restarts_${n}gpu.json 2> gpurun_out/s_restarts_${n}gpu.err; echo "restarts$n rc=$?"; head -c 600 gpurun_out/s_restarts_${n}gpu.json; tail -2 gpurun_out/s_restarts_${n}gpu.err | cut -c1-300;;
  esac
done
