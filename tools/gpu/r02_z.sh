#!/bin/bash
# Round 2, GPU call Z (1 GPU), final: the records committed under profiles/: suite, every-op timing + ncu table, headline ncu (full,
# with source), FFMA2 micro-benchmark, config 5, bench N=1, reference arm, launch list of bench.py.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && timeout 120 /tmp/ubench_ffma2 > gpurun_out/z_ubench_ffma2.jsonl 2>&1
timeout 900 python tools/bench_ops.py > gpurun_out/z_ops.jsonl 2> gpurun_out/z_ops.err
timeout 900 ncu --nvtx --nvtx-include "measure/" --clock-control none \
    --section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --csv --page raw --log-file gpurun_out/z_ops_ncu.csv python tools/bench_ops.py --once --big 8192 --only "^(?!flatten mode)(?!.*diagnosis)(?!.*\[UW)" > gpurun_out/z_ops_ncu.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measure/" -o gpurun_out/z_headline \
    python tools/bench_ops.py --once --only "flatten 16L modes 0-15$|gaussian s20 fast .default" > gpurun_out/z_headline_ncu.out 2>&1
timeout 600 python tools/bench_configs.py --config 3,5 > gpurun_out/z_configs.jsonl 2> gpurun_out/z_configs.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo "bench rc=$?" >> gpurun_out/z_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/z_launches.out 2>&1
tail -4 gpurun_out/z_pytest.log; tail -2 gpurun_out/z_bench_n1.err; cut -c1-300 gpurun_out/z_bench_n1.json
