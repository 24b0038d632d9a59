#!/bin/bash
# Round 2, GPU call U (1 GPU): V-pass ring chunk size; the new peer tests on one GPU; gaussian parity
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -k "gauss or peer or sharpen or blur" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
timeout 600 python tools/bench_ops.py --only "gaussian s20" > gpurun_out/u_gauss.jsonl 2> gpurun_out/u.err
for cg in 1 2 3; do PFE_GAUSS_VCHUNK=$cg timeout 300 python tools/bench_ops.py --only "gaussian s20 (fast \(default\)|EXACT)|gaussian s50" >> gpurun_out/u_gauss_cg$cg.jsonl 2>> gpurun_out/u.err; done
tail -5 gpurun_out/u_pytest.log; tail -3 gpurun_out/u.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/u_gauss*.jsonl')):
    print(f)
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
