#!/bin/bash
# Round 2, GPU call Q (2 GPUs): edge-first strong leg: NCCL parity test + bench N=2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/q_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest_dist.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/q_bench_n2.json 2> gpurun_out/q_bench_n2.err; echo "bench rc=$?" >> gpurun_out/q_bench_n2.err
tail -12 gpurun_out/q_pytest_dist.log; tail -3 gpurun_out/q_bench_n2.err; python - <<PY
import json
d = json.load(open('gpurun_out/q_bench_n2.json'))
print(d["value"], d["ms_per_step"]); print(json.dumps(d["strong"])); print(json.dumps(d["config4"]))
PY
