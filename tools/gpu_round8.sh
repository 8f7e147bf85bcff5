#!/bin/bash
# Round-1 record run: GPU tests, full bench line, ncu launch list of the bench command, ncu --set full of the ALD kernel.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench (default flags)"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-1500
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --levels 16 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/launches.csv | cut -c1-200
echo "== ncu full (ALD kernel, 16 levels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 1 -c 1 -o gpurun_out/prof_ald -f python bench.py --steps 1 --warmup 1 --levels 16 --no-cpu-baseline > gpurun_out/ncu_ald.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_ald.log
ls -la gpurun_out
