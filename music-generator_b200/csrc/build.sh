#!/bin/bash
# Builds libdeepj_sm100.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libdeepj_sm100.so
SRCS="api.cu frontend.cu gemm_simt.cu gemm_tc.cu lstm_scan.cu lstm_scan_tc.cu head_bwd.cu generate.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p ../build
pids=()
for s in $SRCS; do
  o=../build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ dj_common.cuh -nt "$o" ] || [ dj_tc.cuh -nt "$o" ] || [ ../../include/deepj_b200.h -nt "$o" ]; then
    nvcc $FLAGS ${DJ_NVCC_EXTRA} -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for s in $SRCS; do OBJS="$OBJS ../build/${s%.cu}.o"; done
nvcc -shared -o $OUT $OBJS -cudart static
echo "built $(realpath $OUT)"
