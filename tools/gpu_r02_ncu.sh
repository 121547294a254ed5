#!/bin/bash
# round 2: ncu evidence.  bash tools/gpu_r02_ncu.sh train|gen   (one gpurun call each: gpurun_out/ is capped at 64 MiB)
# Each ncu run follows a plain run of the same command line that exited 0.
mkdir -p gpurun_out
export DJ_GRAPH=0      # launch list / captures of the individually launched kernels (a graph replays the same kernels)
TRAIN="python bench.py --steps 3 --warmup 3 --no-generation --no-cpu-baseline --no-kernel-table"
GEN="python bench.py --workload gen1024 --steps 2 --warmup 3 --no-cpu-baseline"
if [ "$1" = "train" ]; then
  $TRAIN > gpurun_out/r02_ncu_plain_train.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 120 --csv --log-file gpurun_out/r02_launches_train.csv $TRAIN > gpurun_out/r02_ncu_launches_train.log 2>&1
  echo "launch list rc=$?"
  $TRAIN > gpurun_out/r02_ncu_plain_train2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"scan_tc_bwd_kernel|scan_tc_fwd_kernel" -s 24 -c 8 -o gpurun_out/r02_prof_scans $TRAIN > gpurun_out/r02_ncu_full_scans.log 2>&1
  echo "full scans rc=$?"
  $TRAIN > gpurun_out/r02_ncu_plain_train3.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"gate_gemm_kernel|layer_input_kernel" -s 33 -c 11 -o gpurun_out/r02_prof_gemm $TRAIN > gpurun_out/r02_ncu_full_gemm.log 2>&1
  echo "full gemm rc=$?"
else
  $GEN > gpurun_out/r02_ncu_plain_gen.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 72 --csv --log-file gpurun_out/r02_launches_gen.csv $GEN > gpurun_out/r02_ncu_launches_gen.log 2>&1
  echo "gen launch list rc=$?"
  $GEN > gpurun_out/r02_ncu_plain_gen2.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"scan_tc_fwd_kernel|gate_gemm_kernel|gen_sample_kernel|frontend_fwd_kernel" -s 52 -c 7 -o gpurun_out/r02_prof_gen $GEN > gpurun_out/r02_ncu_full_gen.log 2>&1
  echo "full gen rc=$?"
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv; du -sh gpurun_out
