#!/bin/bash
# Round-end measurement set: gpu tests, smoke, bench (both arms), ncu launch list, ncu --set full of the
# dominant kernels.  Every ncu command is preceded by the same command run plain.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')}, d['e2e']['value'], d['generation'], d['clocks'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['cpu_baseline'])
PY
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 110 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
[ -n "$SKIP_FULL" ] && exit 0
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_bwd_kernel -s 14 -c 1 -o gpurun_out/prof_scan_tc_bwd_time $BENCH > gpurun_out/ncu_full1.log 2>&1
echo "ncu bwd rc=$?"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_fwd_kernel -s 13 -c 1 -o gpurun_out/prof_scan_tc_fwd_time $BENCH > gpurun_out/ncu_full2.log 2>&1
echo "ncu fwd rc=$?"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gate_gemm_kernel -s 25 -c 1 -o gpurun_out/prof_gate_gemm $BENCH > gpurun_out/ncu_full3.log 2>&1
echo "ncu gemm rc=$?"
