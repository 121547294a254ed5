#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "scan" --timeout 300 -x 2>&1 | grep -vE "^E   " | tail -15
python -m pytest tests/test_gpu_model.py -q -m gpu -k "scaled or bf16" --timeout 600 2>&1 | grep -vE "^E    " | tail -25
