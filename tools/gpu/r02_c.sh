#!/bin/bash
# Round 2, GPU call C (1 GPU): FFMA2 operand-form micro-benchmark; flatten CTA shapes.
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && timeout 120 /tmp/ubench_ffma2 > gpurun_out/c_ubench_ffma2.jsonl 2>&1
for shape in 256 448 512 320; do
  PFE_FLATTEN_BLOCK=$shape timeout 300 python tools/bench_ops.py --only "flatten 16L" > gpurun_out/c_flatten_$shape.jsonl 2>> gpurun_out/c.err
done
cat gpurun_out/c_ubench_ffma2.jsonl; cat gpurun_out/c_flatten_*.jsonl | cut -c1-120
