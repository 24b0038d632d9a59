#!/bin/bash
# Round 2, GPU call L (1 GPU): flatten with branch-free VividLight / table SoftLight: parity + per-mode timing.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blend or flatten or stack or tile or golden or headline or 8k" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
timeout 300 python tools/bench_ops.py --only "flatten" > gpurun_out/l_flatten.jsonl 2> gpurun_out/l.err
tail -4 gpurun_out/l_pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/l_flatten.jsonl'):
    d = json.loads(l); print(f"{d['op']:45s} {d['ms']:.4f}")
PY
