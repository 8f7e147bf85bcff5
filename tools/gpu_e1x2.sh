#!/bin/bash
# engine 1 with two CTAs per SM: parity tests, short bench, op profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --levels 24 --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_short.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_short.log | cut -c1-600
timeout 300 python tools/profile_ops.py 296 tf32x3 > gpurun_out/ops_profile_e1x2.txt 2>&1; head -45 gpurun_out/ops_profile_e1x2.txt
