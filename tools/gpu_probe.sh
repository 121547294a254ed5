#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "scan" --timeout 200 -x 2>&1 | grep -vE "^E  " | tail -25
python tools/scan_probe.py 16 && python tools/scan_probe.py 64
python -m pytest tests/test_gpu_model.py -q -m gpu -k "bf16" --timeout 300 2>&1 | tail -3
