#!/bin/bash
# Round 2, GPU call M (1 GPU): flatten after the mode changes: parity, timing, ncu source-level capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "blend or flatten or stack or tile or golden or headline or 8k" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
timeout 300 python tools/bench_ops.py --only "flatten" > gpurun_out/m_flatten.jsonl 2> gpurun_out/m.err
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measure/" -o gpurun_out/m_flatten \
    python tools/bench_ops.py --once --only "flatten 16L modes 0-15$" > gpurun_out/m_ncu.out 2>&1
tail -4 gpurun_out/m_pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/m_flatten.jsonl'):
    d = json.loads(l); print(f"{d['op']:45s} {d['ms']:.4f}")
PY
