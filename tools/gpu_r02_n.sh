#!/bin/bash
# round 2, call N: the whole GPU suite and the 1-GPU bench lines of the final build
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02n_pytest.log 2>&1
echo "pytest gpu exit $?"; tail -n 4 gpurun_out/r02n_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r02n_smoke.log
timeout 900 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench default exit $?"
timeout 600 python bench.py --workload gen1024 --no-cpu-baseline > gpurun_out/r02n_bench_gen1024.json 2> gpurun_out/r02n_bench_gen1024.err; echo "bench gen1024 exit $?"
timeout 600 python bench.py --workload gen1 > gpurun_out/r02n_bench_gen1.json 2> gpurun_out/r02n_bench_gen1.err; echo "bench gen1 exit $?"
python - <<'PY'
import json
for f in ("r02n_bench", "r02n_bench_gen1024", "r02n_bench_gen1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), "; clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
        g = d.get("generation")
        if g: print("   generation:", g["timesteps_per_s"], "batched", g["batched"]["timesteps_per_s"], "cpu", g["cpu_baseline"]["value"], "equal", g["events_equal_oracle"])
    except Exception as e:
        print(f, "unreadable", e)
PY
