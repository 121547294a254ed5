#!/bin/bash
# round 2, call Q: third backward stream (style / conv gradients)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "train or deterministic or graph or fit or trajectory" --timeout 900 > gpurun_out/r02q_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r02q_pytest.log
for s in 3 2 3 2; do
  DJ_BWD_STREAMS=$s timeout 600 python bench.py --steps 20 --no-generation --no-cpu-baseline --no-kernel-table > gpurun_out/r02q_bench_s$s.json 2> gpurun_out/r02q_bench_s$s.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02q_bench_s$s.json").read().strip().splitlines()[-1])
print("streams $s:", round(d["value"]), "seqs/s", round(d["ms_per_step"], 3), "ms; eager", round(d["ms_per_step_launched_from_python"], 3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
PY
done
