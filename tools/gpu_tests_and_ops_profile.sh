#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for p in tf32x3 tf32; do echo "== op profile $p"; timeout 300 python tools/profile_ops.py 148 $p > gpurun_out/ops_profile_$p.txt 2>&1; echo "rc=$?"; head -24 gpurun_out/ops_profile_$p.txt; done
