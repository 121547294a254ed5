#!/bin/bash
# round 2, final build: ncu evidence for the 1-sequence generation path and the training launch list.
# Each ncu run follows a plain run of the same command line that exited 0.
mkdir -p gpurun_out
export DJ_GRAPH=0
GEN="python bench.py --workload gen1 --steps 4 --warmup 3 --no-cpu-baseline"
TRAIN="python bench.py --steps 3 --warmup 3 --no-generation --no-cpu-baseline --no-kernel-table"
$GEN > gpurun_out/r02f_ncu_plain_gen1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches_gen1.csv $GEN > gpurun_out/r02f_ncu_launches_gen1.log 2>&1
echo "gen1 launch list rc=$?"
$GEN > gpurun_out/r02f_ncu_plain_gen1b.log 2>&1 &&
ncu --set full --clock-control none -k regex:"scan_tc_gen2_kernel|gen_sample1_kernel" -s 8 -c 4 -o gpurun_out/r02f_prof_gen1 $GEN > gpurun_out/r02f_ncu_full_gen1.log 2>&1
echo "gen1 full rc=$?"
$TRAIN > gpurun_out/r02f_ncu_plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 120 --csv --log-file gpurun_out/r02f_launches_train.csv $TRAIN > gpurun_out/r02f_ncu_launches_train.log 2>&1
echo "train launch list rc=$?"
ls -la gpurun_out/r02f_prof_gen1.ncu-rep gpurun_out/r02f_launches_*.csv
