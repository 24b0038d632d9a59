#!/bin/bash
# Round 2, GPU call E (1 GPU): full GPU suite, Gaussian after the broadcast-weight change (timing + ncu), bench N=1.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian|sharpen" > gpurun_out/e_gauss.jsonl 2> gpurun_out/e_gauss.err
timeout 600 ncu --nvtx --nvtx-include "measure/" --clock-control none \
    --section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --csv --page raw --log-file gpurun_out/e_gauss_ncu.csv python tools/bench_ops.py --once --only "gaussian s20 fast .default|gaussian s20 EXACT|gaussian s50|gaussian s4" > gpurun_out/e_gauss_ncu.out 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err; echo "bench rc=$?" >> gpurun_out/e_bench_n1.err
tail -3 gpurun_out/e_pytest.log; cut -c1-100 gpurun_out/e_gauss.jsonl; tail -3 gpurun_out/e_bench_n1.err; cut -c1-600 gpurun_out/e_bench_n1.json
