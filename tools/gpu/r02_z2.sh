#!/bin/bash
# Round 2, GPU call Z2 (1 GPU), final: suite, every-op timing + ncu table (median with 8-bit column counters included)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
timeout 900 python tools/bench_ops.py > gpurun_out/z_ops.jsonl 2> gpurun_out/z_ops.err
timeout 900 ncu --nvtx --nvtx-include "measure/" --clock-control none \
    --section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --csv --page raw --log-file gpurun_out/z_ops_ncu.csv python tools/bench_ops.py --once --big 8192 --only "^(?!flatten mode)(?!.*diagnosis)(?!.*\[)" > gpurun_out/z_ops_ncu.out 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1
tail -4 gpurun_out/z_pytest.log; tail -2 gpurun_out/z_smoke.log; tail -2 gpurun_out/z_ops.err
