#!/bin/bash
# round 2, call H: CTA-pair reverse scans (tcgen05 cta_group::2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "scan_bwd_tcgen05" --timeout 300 > gpurun_out/r02h_pytest_bwd.log 2>&1
echo "pytest bwd exit $?"; tail -n 15 gpurun_out/r02h_pytest_bwd.log
for pair in 1 0; do
  DJ_BWD_PAIR=$pair DJ_DEBUG_OCC=1 timeout 300 python tools/scan_probe.py 64 bwd bf16 > gpurun_out/r02h_probe_pair$pair.log 2>&1
  echo "probe pair=$pair exit $?"; grep -E "tc_bwd|resident" gpurun_out/r02h_probe_pair$pair.log | sort | uniq -c | tail -8
done
