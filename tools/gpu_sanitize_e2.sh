#!/bin/bash
# compute-sanitizer over the tcgen05 engine (tiny run, SBC2_S=2 so that a group stacks two samples): memcheck, racecheck, synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool (engine 2)"
  SBC2_S=2 timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=sbc2_ald_kernel --print-limit 20 python tools/sanitize_run.py fp16x2 > gpurun_out/sanitizer_e2_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run|Error|hazard|Invalid" gpurun_out/sanitizer_e2_$tool.log | head -12
done
