#!/bin/bash
# full GPU suite the way the driver runs it + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')}, d['e2e']['value'], d['generation'], d['clocks'], d['roofline']['frac'], d['cpu_baseline'])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
