#!/bin/bash
# Round 2, GPU call G (1 GPU): full parity suite (new kernels + pipeline + CLI), Gaussian H/V diagnosis, config 5 e2e.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian s20 (H|V) pass|gaussian s20 fast .default" > gpurun_out/g_gauss_hv.jsonl 2> gpurun_out/g.err
timeout 600 python tools/bench_configs.py --config 5 > gpurun_out/g_config5.jsonl 2>> gpurun_out/g.err
tail -15 gpurun_out/g_pytest.log; cut -c1-140 gpurun_out/g_gauss_hv.jsonl; cat gpurun_out/g_config5.jsonl | cut -c1-700; tail -5 gpurun_out/g.err
