#!/bin/bash
# N-GPU check of the fused peer-memory exchange: tests, NCCL-vs-peer equivalence + timing, bench with either exchange.
#   gpurun --gpus 2 -- 'bash tools/gpu_peer.sh 2'
N=${1:-2}
mkdir -p gpurun_out
export DJ_PEER_TIMEOUT_MS=${DJ_PEER_TIMEOUT_MS:-2000}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_gpu_peer.py -x -q -m gpu -k "not two_ranks" > gpurun_out/peer_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/peer_tests.log
if [ -z "$SKIP_CHECK" ]; then DJ_PEER_DEBUG=1 timeout ${CHECK_TIMEOUT:-150} $TR --master-port 29741 tools/peer_check.py --steps 3 > gpurun_out/peer_check.log 2>&1; rc=$?; else rc=0; fi; echo "check rc=$rc"
grep '^{\|^\[' gpurun_out/peer_check.log | tail -40
[ $rc -eq 0 ] || { tail -40 gpurun_out/peer_check.log; exit 0; }
if [ -z "$SKIP_NCCL" ]; then
DJ_PEER_NADAM=0 timeout ${BENCH_TIMEOUT:-200} $TR --master-port 29742 bench.py --gpus $N --steps 20 --warmup 3 --no-generation > gpurun_out/bench_${N}gpu_nccl.log 2>&1; echo "nccl rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_nccl.log | tail -1 > gpurun_out/bench_${N}gpu_nccl.json; python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu_nccl.json'));print('nccl',d['value'],d['e2e']['value'],d['loss'])"
fi
DJ_PEER_NADAM=1 timeout ${BENCH_TIMEOUT:-200} $TR --master-port 29743 bench.py --gpus $N --steps 20 --warmup 3 --no-generation > gpurun_out/bench_${N}gpu_peer.log 2>&1; echo "peer rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_peer.log | tail -1 > gpurun_out/bench_${N}gpu_peer.json; python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu_peer.json'));print('peer',d['value'],d['e2e']['value'],d['loss'])" || tail -30 gpurun_out/bench_${N}gpu_peer.log
