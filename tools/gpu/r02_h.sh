#!/bin/bash
# Round 2, GPU call H (1 GPU): Gaussian H prefetch + V chunk-aligned groups: parity, timing.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gauss or sharpen or glow or smoke or headline or 8k" > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian|sharpen" > gpurun_out/h_gauss.jsonl 2> gpurun_out/h.err
tail -4 gpurun_out/h_pytest.log; cut -c1-120 gpurun_out/h_gauss.jsonl; tail -3 gpurun_out/h.err
