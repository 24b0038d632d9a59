#!/bin/bash
# Round 2, GPU call S (2 GPUs): halo rows over peer memory: new parity tests + bench N=2 (peer vs NCCL transport)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py "tests/test_gpu_parity.py::test_flatten_peer_stores_twice_and_flags" -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --no-config4 > gpurun_out/s_bench_n2.json 2> gpurun_out/s_bench_n2.err; echo "bench rc=$?" >> gpurun_out/s_bench_n2.err
tail -25 gpurun_out/s_pytest.log; tail -5 gpurun_out/s_bench_n2.err; python - <<PY
import json
d = json.load(open('gpurun_out/s_bench_n2.json'))
print(d["value"], d["ms_per_step"]); print(json.dumps(d["strong"]))
PY
