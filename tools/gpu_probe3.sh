#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_probe.py 64
python tools/gemm_probe.py 64 > gpurun_out/gemm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gate_gemm_kernel -s 12 -c 1 -o gpurun_out/prof_gate_gemm python tools/gemm_probe.py 64 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"
