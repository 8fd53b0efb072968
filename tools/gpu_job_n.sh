#!/bin/bash
mkdir -p gpurun_out
for cfg in "16384 512" "16384 1024"; do for blk in 40 5000; do
PB_OZ_TIMING=$blk timeout 120 python tools/oz_timeline.py $cfg 2>&1 | tail -1
PB_OZ_RED=1 PB_OZ_TIMING=$blk timeout 120 python tools/oz_timeline.py $cfg 2>&1 | tail -1
done; done | tee gpurun_out/n_timeline.log
