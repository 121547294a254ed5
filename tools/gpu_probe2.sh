#!/bin/bash
mkdir -p gpurun_out
python tools/scan_probe.py 64 bwd > gpurun_out/probe_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_scan_tc_bwd_time python tools/scan_probe.py 64 bwd > gpurun_out/ncu_tc2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc2.log
python tools/scan_probe.py 64 fwd > gpurun_out/probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_scan_tc_fwd_time python tools/scan_probe.py 64 fwd > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc.log
python -m pytest tests/test_gpu_model.py -q -m gpu -k "chunks_of_32 or bf16 or golden" --timeout 300 2>&1 | tail -3
python tools/quick_bench.py 64 2>&1 | grep -A14 "^bf16" | head -16
