#!/bin/bash
# Lean-kernel check: GPU tests, per-op cycle profiles (both precisions), short bench.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for p in tf32x3 tf32; do echo "== op profile $p"; timeout 300 python tools/profile_ops.py 148 $p > gpurun_out/ops_profile_$p.txt 2>&1; echo "rc=$?"; head -30 gpurun_out/ops_profile_$p.txt; done
echo "== bench short"; timeout 600 python bench.py --levels 24 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_short.log | cut -c1-300
