#!/bin/bash
python -m pytest tests/test_gpu_model.py -q -m gpu -k "generation or keras_like" --timeout 300 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/quick_bench.py 16 2>&1 | grep -A12 "^generation"
