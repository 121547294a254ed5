#!/bin/bash
# ncu --set full (with source) of the time-axis tensor-core scan kernels at the B=64 training shape
mkdir -p gpurun_out
export DJ_FWD_NS=${DJ_FWD_NS:-1}
python tools/scan_probe.py 64 all > gpurun_out/probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_scan_fwd_time_r01c python tools/scan_probe.py 64 all > gpurun_out/ncu_scans.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_scan_bwd_time_r01c python tools/scan_probe.py 64 all >> gpurun_out/ncu_scans.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scans.log; cat gpurun_out/probe_plain.log
