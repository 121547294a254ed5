#!/bin/bash
# round 2, call B: full parity suite (one process per file), scan probe, bench
mkdir -p gpurun_out
for f in kernels model peer; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02b_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02b_pytest_$f.log
  grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/r02b_pytest_$f.log | tail -8
done
timeout 300 python tools/scan_probe.py 64 > gpurun_out/r02b_scan_probe.log 2>&1
DJ_BWD_SHARED=0 timeout 300 python tools/scan_probe.py 64 bwd bf16 >> gpurun_out/r02b_scan_probe.log 2>&1
cat gpurun_out/r02b_scan_probe.log
timeout 600 python bench.py --steps 10 --no-generation --no-cpu-baseline > gpurun_out/r02b_bench_mixed.json 2> gpurun_out/r02b_bench_mixed.err
DJ_BWD_SHARED=0 timeout 600 python bench.py --steps 10 --no-generation --no-cpu-baseline > gpurun_out/r02b_bench_mixed_2wave.json 2> gpurun_out/r02b_bench_mixed_2wave.err
python - <<'PY'
import json
for f in ("gpurun_out/r02b_bench_mixed.json", "gpurun_out/r02b_bench_mixed_2wave.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roof", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["step_roofline"]["frac"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "unreadable", e)
PY
