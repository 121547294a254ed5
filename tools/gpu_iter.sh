#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu --timeout 600 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['loss'], d['roofline']['avg_launch_ms'], d['clocks'])"; done
