#!/bin/bash
# Round 2, GPU call SAN (1 GPU): compute-sanitizer over tools/sanitize_smoke.py (every kernel family incl. the round-2 ones)
set -x
mkdir -p gpurun_out
timeout 120 python tools/sanitize_smoke.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/san_plain.log; tail -3 gpurun_out/san_plain.log
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/san_$tool.log
  grep -c "max diff" gpurun_out/san_$tool.log; grep "SANITIZE_SMOKE\|ERROR SUMMARY\|RACECHECK SUMMARY\|rc=" gpurun_out/san_$tool.log
done
