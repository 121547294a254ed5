#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')}, d['e2e'], d['generation'], d['clocks'])"
tail -3 gpurun_out/bench.err
python tools/quick_bench.py 64 2>&1 | grep -A45 "^bf16" | head -50
