#!/bin/bash
# round 2, call E: CUDA-graph training step
mkdir -p gpurun_out
for f in kernels model peer; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02e_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02e_pytest_$f.log
  grep -E "passed|failed|exit" gpurun_out/r02e_pytest_$f.log | tail -3
  grep -E "^E  |graph capture" gpurun_out/r02e_pytest_$f.log | head -20
done
timeout 600 python bench.py --steps 20 --no-generation --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
timeout 600 python bench.py --steps 20 --batch 16 --no-generation --no-cpu-baseline > gpurun_out/r02e_bench_b16.json 2> gpurun_out/r02e_bench_b16.err
timeout 600 python bench.py --steps 20 --batch 8 --no-generation --no-cpu-baseline > gpurun_out/r02e_bench_b8.json 2> gpurun_out/r02e_bench_b8.err
DJ_GRAPH=0 timeout 600 python bench.py --steps 20 --batch 8 --no-generation --no-cpu-baseline > gpurun_out/r02e_bench_b8_nograph.json 2> gpurun_out/r02e_bench_b8_nograph.err
python - <<'PY'
import json
for f in ("r02e_bench", "r02e_bench_b16", "r02e_bench_b8", "r02e_bench_b8_nograph"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "eager", round(d["ms_per_step_launched_from_python"], 3), "graph", d["cuda_graph"], "e2e", round(d["e2e"]["value"]), "host enq", round(d["host_enqueue_ms_per_step"], 3), d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], "loss", d["loss"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r02e_bench.err
