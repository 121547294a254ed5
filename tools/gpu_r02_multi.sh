#!/bin/bash
# round 2: multi-GPU measurements on N GPUs of one box:  bash tools/gpu_r02_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, port, args...
  name=$1; port=$2; shift 2
  timeout 900 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r02_${name}_${N}gpu.json 2> gpurun_out/r02_${name}_${N}gpu.err
  echo "$name rc=$?"; tail -c 600 gpurun_out/r02_${name}_${N}gpu.json | head -c 600; echo
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader | head -8
run weak 29511 --steps 20 --warmup 5 --no-kernel-table
run strong 29512 --steps 20 --warmup 5 --scaling strong --no-kernel-table
run gen1024 29513 --workload gen1024 --steps 8
run scaled 29514 --workload scaled --steps 5 --no-kernel-table
run ref 29515 --impl reference --steps 3 --warmup 1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest_peer_2gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_peer_2gpu.log
fi
python - <<PY
import json
for name in ("weak", "strong", "gen1024", "scaled", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/r02_{name}_${N}gpu.json").read().strip().splitlines()[-1])
        print(name, "N=$N", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("rank_ms_per_step"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(name, "unreadable", e)
PY
