#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -q -m gpu -k "odd_shape" --timeout 300 2>&1 | tail -3
python generate.py --bars 1 --styles 0 5 12 2>&1 | tail -6; ls -la out/samples/ 2>/dev/null
python train.py --epochs 3 --num-seqs 48 2>&1 | tail -8; ls -la out/ 2>/dev/null
python - <<'PY'
import numpy as np, model
m = model.build_models(precision="fp32")
m[0].load_weights("out/model.h5")
print("reloaded weights ok; style layer", m[0].get_layer("style").get_weights()[0].shape, "params", m[0].count_params())
PY
