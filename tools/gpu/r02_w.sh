#!/bin/bash
# Round 2, GPU call W (1 GPU): V pass with the next chunk awaited before the current chunk's last group; chunk size sweep
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gauss or sharpen or blur" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
timeout 600 python tools/bench_ops.py --only "gaussian s20 (V pass|fast \(default\)|EXACT)|gaussian s50" > gpurun_out/w_gauss.jsonl 2> gpurun_out/w.err
for cg in 2 3 9; do PFE_GAUSS_VCHUNK=$cg timeout 300 python tools/bench_ops.py --only "gaussian s20 (fast \(default\)|EXACT)|gaussian s50" >> gpurun_out/w_gauss_cg$cg.jsonl 2>> gpurun_out/w.err; done
tail -3 gpurun_out/w_pytest.log; tail -3 gpurun_out/w.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/w_gauss*.jsonl')):
    print(f)
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if 'ms' in d: print('  ', d['op'], round(d['ms'],4))
PY
