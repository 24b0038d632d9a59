#!/bin/bash
# Round 2, GPU call F (1 GPU): new median / HSL / vignette / motion / brush kernels: parity suite, op timings;
# Gaussian inner-loop micro-benchmark (per-step LDS / LDCU variants).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && timeout 120 /tmp/ubench_ffma2 > gpurun_out/f_ubench_ffma2.jsonl 2>&1
timeout 600 python tools/bench_ops.py --only "median|adjust|vignette|motion|brush|box" > gpurun_out/f_ops.jsonl 2> gpurun_out/f_ops.err
PFE_MEDIAN_KERNEL=bisect timeout 300 python tools/bench_ops.py --only "median r2|median r7" > gpurun_out/f_ops_bisect.jsonl 2>> gpurun_out/f_ops.err
tail -5 gpurun_out/f_pytest.log; grep -E "V[5-8]" gpurun_out/f_ubench_ffma2.jsonl; cut -c1-110 gpurun_out/f_ops.jsonl gpurun_out/f_ops_bisect.jsonl
