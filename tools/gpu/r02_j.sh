#!/bin/bash
# Round 2, GPU call J (1 GPU): Gaussian with the carried first input: parity + timing; full suite.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian|sharpen" > gpurun_out/j_gauss.jsonl 2> gpurun_out/j.err
tail -4 gpurun_out/j_pytest.log; cut -c1-120 gpurun_out/j_gauss.jsonl; tail -3 gpurun_out/j.err
