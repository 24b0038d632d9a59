#!/bin/bash
# Round 2, last GPU call (1 GPU): what the driver runs at round end - suite, smoke, reference arm, bench with defaults
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" >> gpurun_out/final_bench.err
tail -3 gpurun_out/final_pytest.log; tail -1 gpurun_out/final_smoke.log; tail -5 gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench_reference.json
