#!/bin/bash
# First-light run on a B200 box: smoke (both parameter paths), GPU tests, short bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke, parameters from L2 (no staging)"; SBC_STAGE_WEIGHTS=0 timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_nostage.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke_nostage.log
echo "== smoke, staged"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench short"; timeout 600 python bench.py --levels 48 --steps 2 --warmup 1 > gpurun_out/bench_short.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_short.log
