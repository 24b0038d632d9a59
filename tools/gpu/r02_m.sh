#!/bin/bash
# Round 2, GPU call M (1 GPU): median with 8-bit column counters: parity, timing against the 16-bit layout
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "median" > gpurun_out/m_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/m_pytest.log
tail -5 gpurun_out/m_pytest.log
timeout 600 python tools/bench_ops.py --only "median" > gpurun_out/m_ops.jsonl 2> gpurun_out/m.err
tail -3 gpurun_out/m.err
python - <<PY
import json
for l in open('gpurun_out/m_ops.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
