#!/bin/bash
# Builds probit_b200/libprobit_b200.so for sm_100a (in-tree, so it travels with gpurun snapshots).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libprobit_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
mkdir -p ../../build/obj
OBJS=""
for f in capi gemm_dmma potrf likelihood gram blas2 fit; do
  $NVCC $FLAGS -c $f.cu -o ../../build/obj/$f.o 2> ../../build/obj/$f.ptxas.log || { cat ../../build/obj/$f.ptxas.log; exit 1; }
  OBJS="$OBJS ../../build/obj/$f.o"
done
$NVCC -shared -o $OUT $OBJS -cudart shared
echo "built $OUT"
