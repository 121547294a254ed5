#!/bin/bash
# round 2, call D: full parity suite, scan probe, bench (half-precision saved gates, 16-sequence inference tiles, one-round mask hash)
mkdir -p gpurun_out
for f in kernels model peer; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02d_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02d_pytest_$f.log
  grep -E "passed|failed|exit" gpurun_out/r02d_pytest_$f.log | tail -3
  grep -E "north-star|worst grad|lock-step|scan_tc_infer" gpurun_out/r02d_pytest_$f.log | head -30
done
timeout 300 python tools/scan_probe.py 64 > gpurun_out/r02d_scan_probe.log 2>&1; cat gpurun_out/r02d_scan_probe.log
timeout 600 python bench.py --steps 10 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 600 python bench.py --steps 10 --batch 16 --no-generation --no-cpu-baseline > gpurun_out/r02d_bench_b16.json 2> gpurun_out/r02d_bench_b16.err
timeout 600 python bench.py --steps 10 --precision bf16 --no-generation --no-cpu-baseline > gpurun_out/r02d_bench_bf16.json 2> gpurun_out/r02d_bench_bf16.err
python - <<'PY'
import json
for f in ("r02d_bench", "r02d_bench_b16", "r02d_bench_bf16"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "host enq", round(d["host_enqueue_ms_per_step"], 3), d["clocks"]["sm_mhz"], "roof", round(d["roofline"]["frac"], 3), round(d["step_roofline"]["frac"], 3))
        if d.get("generation"): print("   gen", d["generation"])
        if d.get("kernels"): print("   ", {k: v["ms_per_step"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
