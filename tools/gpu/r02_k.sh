#!/bin/bash
# Round 2, GPU call K (1 GPU): full suite after the asynchronous masked Gaussian, warp_region and copy-on-write tiles.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/k_smoke.log
tail -25 gpurun_out/k_pytest.log; tail -3 gpurun_out/k_smoke.log
