#!/bin/bash
# compute-sanitizer over the fused kernel (tiny run): memcheck, racecheck (shared-memory hazards), synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --kernel-regex kns=sbc_ald_kernel --print-limit 20 python tools/sanitize_run.py tf32x3 > gpurun_out/sanitizer_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
