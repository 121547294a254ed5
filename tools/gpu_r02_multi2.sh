#!/bin/bash
# round 2, final build: multi-GPU lines on N GPUs of one box:  bash tools/gpu_r02_multi2.sh N [weak strong gen1024 ...]
N=${1:-2}; shift
WHAT=${@:-weak}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, port, args...
  name=$1; port=$2; shift 2
  timeout 900 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r02f_${name}_${N}gpu.json 2> gpurun_out/r02f_${name}_${N}gpu.err
  echo "$name rc=$?"
}
for w in $WHAT; do
  case $w in
    weak) run weak 29511 --steps 20 --warmup 5 --no-kernel-table ;;
    strong) run strong 29512 --steps 20 --warmup 5 --scaling strong --no-kernel-table ;;
    gen1024) run gen1024 29513 --workload gen1024 --steps 8 ;;
    scaled) run scaled 29514 --workload scaled --steps 5 --no-kernel-table ;;
    peer) timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q --timeout 600 > gpurun_out/r02f_pytest_peer_${N}gpu.log 2>&1; tail -3 gpurun_out/r02f_pytest_peer_${N}gpu.log ;;
  esac
done
python - <<PY
import json
for name in "$WHAT".split():
    try:
        d = json.loads(open(f"gpurun_out/r02f_{name}_${N}gpu.json").read().strip().splitlines()[-1])
        print(name, "N=$N", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("rank_ms_per_step"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(name, "unreadable", e)
PY
