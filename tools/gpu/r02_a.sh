#!/bin/bash
# Round 2, GPU call A (1 GPU): parity of the uniform-weight Gaussian kernels, a timing pass over every op,
# the FFMA2 operand micro-benchmark, and an ncu metric pass over one launch of every kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && timeout 60 /tmp/ubench_ffma2 > gpurun_out/a_ubench_ffma2.txt 2>&1
timeout 600 python tools/bench_ops.py > gpurun_out/a_ops.jsonl 2> gpurun_out/a_ops.err
timeout 900 ncu --nvtx --nvtx-include "measure/" --clock-control none \
    --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy \
    --section ComputeWorkloadAnalysis --section SchedulerStats \
    --csv --page raw --log-file gpurun_out/a_ops_ncu.csv python tools/bench_ops.py --once --big 8192 > gpurun_out/a_ops_ncu.out 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_ubench_ffma2.txt; cat gpurun_out/a_ops.jsonl | cut -c1-200
