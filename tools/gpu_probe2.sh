#!/bin/bash
mkdir -p gpurun_out
python tools/scan_probe.py 16 fwd > gpurun_out/probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_scan_tc_fwd2 python tools/scan_probe.py 16 fwd > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc.log
python tools/scan_probe.py 16 bwd > gpurun_out/probe_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_bwd_kernel -s 3 -c 1 -o gpurun_out/prof_scan_tc_bwd python tools/scan_probe.py 16 bwd > gpurun_out/ncu_tc2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc2.log
