#!/bin/bash
# round 2, call C: kernel + model tests, bench with generation on the tensor cores
mkdir -p gpurun_out
for f in kernels model; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rP --timeout 900 > gpurun_out/r02c_pytest_$f.log 2>&1
  echo "pytest $f exit $?" >> gpurun_out/r02c_pytest_$f.log
  grep -E "passed|failed|exit" gpurun_out/r02c_pytest_$f.log | tail -4
  grep -E "^___|half-split GEMM|scan_tc_infer|lock-step" gpurun_out/r02c_pytest_$f.log | head -40
done
timeout 600 python bench.py --steps 10 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
DJ_GEMM_BN=128 timeout 600 python bench.py --steps 10 --no-generation --no-cpu-baseline > gpurun_out/r02c_bench_bn128.json 2> gpurun_out/r02c_bench_bn128.err
DJ_GEN_TC=0 timeout 600 python bench.py --workload gen1024 --steps 8 --no-cpu-baseline > gpurun_out/r02c_gen1024_simt.json 2> gpurun_out/r02c_gen1024_simt.err
timeout 600 python bench.py --workload gen1024 --steps 8 --no-cpu-baseline > gpurun_out/r02c_gen1024_tc.json 2> gpurun_out/r02c_gen1024_tc.err
python - <<'PY'
import json
for f in ("r02c_bench", "r02c_bench_bn128", "r02c_gen1024_simt", "r02c_gen1024_tc"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
        if d.get("generation"): print("   gen", d["generation"])
        if d.get("kernels"): print("   ", {k: v["ms_per_step"] for k, v in d["kernels"].items() if "gemm" in k})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r02c_*.err
