#!/bin/bash
python -m pytest tests/test_gpu_model.py -q -m gpu -k "train or golden or scaled" --timeout 600 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('two-stream', d['value'], d['ms_per_step'], d['e2e']['value'])"
DJ_NO_OVERLAP=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('one-stream', d['value'], d['ms_per_step'], d['e2e']['value'])"
