#!/bin/bash
# Builds probit_b200/libprobit_b200.so for sm_100a (in-tree, so it travels with gpurun snapshots).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${PB_OUT:-../libprobit_b200.so}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -Wno-deprecated-gpu-targets"
OBJ=${PB_OBJ:-../../build/obj}
mkdir -p $OBJ
SRCS="capi gemm_dmma ozaki potrf likelihood gram blas2 fit dist"
pids=""
for f in $SRCS; do
  ( $NVCC $FLAGS -c $f.cu -o $OBJ/$f.o 2> $OBJ/$f.ptxas.log || { cat $OBJ/$f.ptxas.log; exit 1; } ) &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS $OBJ/$f.o"; done
# libnccl is reached through dlopen at run time (dist.cu): no link-time dependency beyond libcudart and libdl
$NVCC -shared -o $OUT $OBJS -cudart shared -ldl
echo "built $OUT"
