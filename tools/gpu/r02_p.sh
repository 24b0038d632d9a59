#!/bin/bash
# Round 2, GPU call P (N GPUs, N = $1): NCCL band-split parity test (2 ranks), then bench.py at N (weak + strong + config 4) and the
# reference arm launched the way the driver launches it.
set -x
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/p_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/p_pytest_dist_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest_dist_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/p_bench_reference_n$N.json 2> gpurun_out/p_bench_reference_n$N.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/p_bench_n$N.json 2> gpurun_out/p_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/p_bench_n$N.err
tail -3 gpurun_out/p_pytest_dist_n$N.log; tail -3 gpurun_out/p_bench_n$N.err; python - <<PY
import json
d = json.load(open('gpurun_out/p_bench_n$N.json'))
print(d["value"], d["ms_per_step"]); print(json.dumps(d["e2e"])); print(json.dumps(d["strong"])); print(json.dumps(d["config4"]))
r = json.load(open('gpurun_out/p_bench_reference_n$N.json')); print(r["value"], r["cpu_baseline"])
PY
