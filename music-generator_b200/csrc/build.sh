#!/bin/bash
# Builds libdeepj_sm100.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=${DJ_OUT:-../libdeepj_sm100.so}
BLD=${DJ_BUILD_DIR:-../build}
SRCS="api.cu frontend.cu gemm_simt.cu gemm_tc.cu lstm_scan.cu lstm_scan_tc.cu head_bwd.cu generate.cu peer_nadam.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p $BLD
pids=()
for s in $SRCS; do
  o=$BLD/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ dj_common.cuh -nt "$o" ] || [ dj_tc.cuh -nt "$o" ] || [ ../../include/deepj_b200.h -nt "$o" ]; then
    nvcc $FLAGS ${DJ_NVCC_EXTRA} -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for s in $SRCS; do OBJS="$OBJS $BLD/${s%.cu}.o"; done
nvcc -shared -o $OUT $OBJS -cudart static
echo "built $(realpath $OUT)"
