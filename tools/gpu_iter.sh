#!/bin/bash
# one development iteration on the GPU: full gpu test-suite, kernel probes, per-kernel timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -4
python tools/scan_probe.py 64 all
python tools/quick_bench.py 64 2>&1 | grep -v "^fp32" | tail -75
