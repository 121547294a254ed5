#!/bin/bash
# round 2, call A: parity suite + scan probe + bench in both precisions
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
for f in kernels model peer; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02a_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02a_pytest_$f.log
  grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/r02a_pytest_$f.log | tail -15
done
timeout 300 python tools/scan_probe.py 64 > gpurun_out/r02a_scan_probe.log 2>&1
cat gpurun_out/r02a_scan_probe.log
timeout 600 python bench.py --steps 10 > gpurun_out/r02a_bench_mixed.json 2> gpurun_out/r02a_bench_mixed.err
timeout 300 python bench.py --steps 10 --precision bf16 --no-generation --no-cpu-baseline > gpurun_out/r02a_bench_bf16.json 2> gpurun_out/r02a_bench_bf16.err
python - <<'PY'
import json
for f in ("gpurun_out/r02a_bench_mixed.json", "gpurun_out/r02a_bench_bf16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["step_roofline"]["frac"], d["clocks"])
        print("   gen", d.get("generation"))
    except Exception as e:
        print(f, "unreadable", e)
PY
