#!/bin/bash
# round 2, final build: ncu --set full of the training scans, the split gate GEMM and layer_input (after a plain run)
mkdir -p gpurun_out
export DJ_GRAPH=0
TRAIN="python bench.py --steps 3 --warmup 3 --no-generation --no-cpu-baseline --no-kernel-table"
$TRAIN > gpurun_out/r02f_ncu_plain_train2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_tc_bwd_kernel|scan_tc_fwd_kernel" -s 24 -c 8 -o gpurun_out/r02f_prof_scans $TRAIN > gpurun_out/r02f_ncu_full_scans.log 2>&1
echo "full scans rc=$?"
$TRAIN > gpurun_out/r02f_ncu_plain_train3.log 2>&1 &&
ncu --set full --clock-control none -k regex:"gate_gemm_kernel|layer_input_fast_kernel|wgrad_gemm_kernel" -s 40 -c 12 -o gpurun_out/r02f_prof_gemm $TRAIN > gpurun_out/r02f_ncu_full_gemm.log 2>&1
echo "full gemm rc=$?"
ls -la gpurun_out/r02f_prof_scans.ncu-rep gpurun_out/r02f_prof_gemm.ncu-rep
