#!/bin/bash
# Round record run (1 GPU): GPU tests, smoke, default bench, reference arm, ncu launch list + full capture, per-op cycle profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench (default flags)"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
echo "== ncu launch list (levels 16)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --levels 16 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu launch list of the DEFAULT bench command"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --no-cpu-baseline > gpurun_out/bench_under_ncu_default.log 2>&1; echo "rc=$?"
echo "== ncu full (ALD kernel, 16 levels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 1 -c 1 -o gpurun_out/prof_ald -f python bench.py --steps 1 --warmup 1 --levels 16 --no-cpu-baseline --no-extra > gpurun_out/ncu_ald.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_ald.log | cut -c1-200
timeout 300 python tools/profile_ops.py 296 tf32x3 > gpurun_out/ops_profile_tf32x3.txt 2>&1; sed -n 1,3p gpurun_out/ops_profile_tf32x3.txt
