#!/bin/bash
# Round 2, GPU call Y (1 GPU): work queues in the persistent Gaussian passes: parity, then timings
set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "gauss or sharpen or blur" > gpurun_out/y_pytest_gauss.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/y_pytest_gauss.log
tail -5 gpurun_out/y_pytest_gauss.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python tools/bench_ops.py --only "gaussian|sharpen" > gpurun_out/y_gauss.jsonl 2> gpurun_out/y.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
tail -4 gpurun_out/y_pytest.log; tail -3 gpurun_out/y.err
python - <<PY
import json
for l in open('gpurun_out/y_gauss.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
