#!/bin/bash
# Round 2, GPU call I (1 GPU): full ncu capture with source-level stall sampling of the two Gaussian passes.
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measure/" -o gpurun_out/i_gauss_hv \
    python tools/bench_ops.py --once --only "gaussian s20 (H|V) pass only$" > gpurun_out/i_ncu.out 2>&1
ls -la gpurun_out/i_gauss_hv.ncu-rep; tail -3 gpurun_out/i_ncu.out
