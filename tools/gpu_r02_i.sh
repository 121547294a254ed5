#!/bin/bash
# round 2, call I: pair reverse scan variants (non-trace timings)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "scan_bwd_tcgen05" --timeout 300 > gpurun_out/r02i_pytest_bwd.log 2>&1
echo "pytest bwd exit $?"; tail -n 3 gpurun_out/r02i_pytest_bwd.log
for rep in 1 2; do
for pair in 1 0; do
  DJ_BWD_PAIR=$pair timeout 300 python tools/scan_probe.py 64 bwd bf16 > gpurun_out/r02i_probe_pair$pair.log 2>&1
  echo "probe direct pair=$pair exit $?"; grep -E "tc_bwd" gpurun_out/r02i_probe_pair$pair.log
done
DJ_PROBE_LIB=libdeepj_alt.so DJ_BWD_PAIR=1 timeout 300 python tools/scan_probe.py 64 bwd bf16 > gpurun_out/r02i_probe_alt.log 2>&1
echo "probe forwarded pair=1 exit $?"; grep -E "tc_bwd" gpurun_out/r02i_probe_alt.log
done
