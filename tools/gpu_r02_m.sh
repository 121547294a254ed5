#!/bin/bash
# round 2, call M: barrier-free single-sequence sampler
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "generation or generate or smoke" --timeout 600 -rP > gpurun_out/r02m_pytest_gen.log 2>&1
echo "pytest gen exit $?"; grep -E "passed|failed|error|lock-step|Error|timed out" gpurun_out/r02m_pytest_gen.log | tail -12
timeout 200 python tools/gen_probe.py 1 > gpurun_out/r02m_probe.log 2>&1
echo "probe exit $?"; head -n 5 gpurun_out/r02m_probe.log
timeout 200 python tools/gen_probe.py 32 > gpurun_out/r02m_probe32.log 2>&1
echo "probe 32 exit $?"; head -n 5 gpurun_out/r02m_probe32.log
DJ_GEN_SAMPLE1=0 timeout 200 python tools/gen_probe.py 32 > gpurun_out/r02m_probe32_old.log 2>&1
echo "probe 32 (old sampler) exit $?"; head -n 5 gpurun_out/r02m_probe32_old.log
timeout 300 python bench.py --workload gen1 --no-cpu-baseline > gpurun_out/r02m_bench_gen1.json 2> gpurun_out/r02m_bench_gen1.err; echo "bench gen1 exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m_bench_gen1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step")}, d.get("e2e"))
PY
