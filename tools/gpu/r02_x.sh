#!/bin/bash
# Round 2, GPU call X (1 GPU): start-delay stagger of the persistent Gaussian passes
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_ops.py --only "gaussian s20 (H pass only$|V pass only$|H pass only \[STAGGER|V pass only \[STAGGER|fast \(default\))" > gpurun_out/x_gauss.jsonl 2> gpurun_out/x.err
tail -3 gpurun_out/x.err
python - <<PY
import json
for l in open('gpurun_out/x_gauss.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
