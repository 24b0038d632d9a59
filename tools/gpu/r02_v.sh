#!/bin/bash
# Round 2, GPU call V (1 GPU): V-pass ring chunk size, coarse end
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_ops.py --only "gaussian s20 V pass" > gpurun_out/v_gauss.jsonl 2> gpurun_out/v.err
for cg in 7 9 14; do PFE_GAUSS_VCHUNK=$cg timeout 300 python tools/bench_ops.py --only "gaussian s20 (fast \(default\)|EXACT)|gaussian s50" >> gpurun_out/v_gauss_cg$cg.jsonl 2>> gpurun_out/v.err; done
tail -3 gpurun_out/v.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/v_gauss*.jsonl')):
    print(f)
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
