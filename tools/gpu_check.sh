#!/bin/bash
# One gpurun call: kernel tests, tcgen05 tests (isolated process: a trap poisons the
# context), model parity tests, quick timing probe.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tcgen05" --timeout 300
run tcgen05 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05" --timeout 200 -x
run smoke python __graft_entry__.py smoke
run model_fp32 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not bf16" --timeout 500
run model_bf16 python -m pytest tests/test_gpu_model.py -q -m gpu -k "bf16" --timeout 300
TAILN=60 run quick_bench python tools/quick_bench.py 16
