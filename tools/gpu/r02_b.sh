#!/bin/bash
# Round 2, GPU call B (1 GPU): Gaussian after the H-grid / V-tile / triangular-group changes, EXACT on packed FFMA2,
# flatten CTA shapes.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian|sharpen" > gpurun_out/b_gauss.jsonl 2> gpurun_out/b_gauss.err
PFE_GAUSS_NO_TRI=1 timeout 300 python tools/bench_ops.py --only "gaussian s20 fast .default|gaussian s4|gaussian s20 EXACT" > gpurun_out/b_gauss_notri.jsonl 2>> gpurun_out/b_gauss.err
for shape in 256 448 512 320; do
  PFE_FLATTEN_BLOCK=$shape timeout 300 python tools/bench_ops.py --only "flatten 16L" > gpurun_out/b_flatten_$shape.jsonl 2>> gpurun_out/b_gauss.err
done
timeout 600 ncu --nvtx --nvtx-include "measure/" --clock-control none \
    --section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --csv --page raw --log-file gpurun_out/b_gauss_ncu.csv python tools/bench_ops.py --once --only "gaussian s20 fast .default|gaussian s20 EXACT|gaussian s50" > gpurun_out/b_gauss_ncu.out 2>&1
tail -3 gpurun_out/b_pytest.log; cat gpurun_out/b_gauss.jsonl gpurun_out/b_gauss_notri.jsonl gpurun_out/b_flatten_*.jsonl | cut -c1-120
