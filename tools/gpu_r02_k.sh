#!/bin/bash
# round 2, call K: what the driver runs at round end -- smoke, default bench (both arms), timed
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1; echo "smoke exit $? $(( $(date +%s) - t0 )) s"; tail -n 4 gpurun_out/r02k_smoke.log
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; echo "bench reference exit $? $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench default exit $? $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02k_bench_ref.json").read().strip().splitlines()[-1])
print("ours", round(d["value"]), "seqs/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"]), "; steps", d["steps"], "warmup", d["warmup"], "; clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "avg_launch_ms")})
print("step_roofline", d["step_roofline"]["frac"], "cpu", d["cpu_baseline"])
print("generation", json.dumps(d["generation"])[:900])
print("reference", r["value"], r["unit"], r["cpu_baseline"])
PY
tail -n 3 gpurun_out/r02k_bench.err
