#!/bin/bash
# Round 2, GPU call T (N GPUs, N = $1): strong leg with halo rows over peer memory, NCCL transport timed beside it
N=${1:-4}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/t_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/t_pytest_dist_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest_dist_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t_bench_n$N.json 2> gpurun_out/t_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/t_bench_n$N.err
tail -6 gpurun_out/t_pytest_dist_n$N.log; tail -5 gpurun_out/t_bench_n$N.err; python - <<PY
import json
d = json.load(open('gpurun_out/t_bench_n$N.json'))
print(d["value"], d["ms_per_step"]); print(json.dumps(d["strong"])); print(json.dumps(d["config4"]))
PY
