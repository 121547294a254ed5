#!/bin/bash
# round 2, call J: reverse scans with compile-time strides, note-axis reverse scan on CTA pairs
mkdir -p gpurun_out
for f in kernels model peer; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02j_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02j_pytest_$f.log
  grep -E "passed|failed|exit" gpurun_out/r02j_pytest_$f.log | tail -3
  grep -E "^E  |graph capture" gpurun_out/r02j_pytest_$f.log | head -20
done
timeout 600 python bench.py --steps 20 --no-generation --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
DJ_BWD_PAIR=0 timeout 600 python bench.py --steps 20 --no-generation --no-cpu-baseline > gpurun_out/r02j_bench_det.json 2> gpurun_out/r02j_bench_det.err
python - <<'PY'
import json
for f in ("r02j_bench", "r02j_bench_det"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "eager", d.get("ms_per_step_launched_from_python"), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
        for k, v in (d.get("kernels") or {}).items(): print("   ", k, v)
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r02j_bench.err gpurun_out/r02j_bench_det.err
