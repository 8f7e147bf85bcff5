#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== op profile"; timeout 300 python tools/profile_ops.py 148 > gpurun_out/ops_profile.txt 2>&1; echo "rc=$?"; head -30 gpurun_out/ops_profile.txt
echo "== ncu full (forward B=148)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 3 -c 1 -o gpurun_out/prof_fwd -f python tools/profile_ops.py 148 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
