#!/bin/bash
# Round 2, GPU call O (1 GPU): after moving the selection-blur skipping into its own instantiations: suite, bench N=1, gaussian timing.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
timeout 300 python tools/bench_ops.py --only "gaussian s20|sharpen" > gpurun_out/o_gauss.jsonl 2> gpurun_out/o.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err; echo "bench rc=$?" >> gpurun_out/o_bench_n1.err
tail -3 gpurun_out/o_pytest.log; cut -c1-110 gpurun_out/o_gauss.jsonl; python - <<'PY'
import json
d = json.load(open('gpurun_out/o_bench_n1.json'))
print(d["value"], d["ms_per_step"], {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()}, d["extra"])
PY
