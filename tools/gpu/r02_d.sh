#!/bin/bash
# Round 2, GPU call D (2 GPUs): NCCL band-split parity test, then bench.py at N=2 (weak + strong + config 4).
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/d_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "bench rc=$?" >> gpurun_out/d_bench_n2.err
tail -5 gpurun_out/d_pytest_dist.log; tail -5 gpurun_out/d_bench_n2.err; cut -c1-3000 gpurun_out/d_bench_n2.json
