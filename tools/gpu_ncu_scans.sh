#!/bin/bash
# ncu --set full (with source) of the four tensor-core scan kernels at the B=64 training shapes
mkdir -p gpurun_out
python tools/scan_probe.py 64 all > gpurun_out/probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_ -s 8 -c 4 -o gpurun_out/prof_scans_r01b python tools/scan_probe.py 64 all > gpurun_out/ncu_scans.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scans.log; cat gpurun_out/probe_plain.log
