#!/bin/bash
# Round 2, GPU call F (1 GPU): flatten with a per-warp loop bound and lane*4 kept in a register: parity, timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -k "flatten or blend or stack or tile or peer or smoke" > gpurun_out/f_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
timeout 600 python tools/bench_ops.py --only "flatten" > gpurun_out/f_ops.jsonl 2> gpurun_out/f.err
tail -3 gpurun_out/f.err
python - <<PY
import json
for l in open('gpurun_out/f_ops.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
