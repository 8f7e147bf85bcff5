#!/bin/bash
# Round record run (1 GPU): GPU tests, default bench, reference arm, ncu launch list + full capture of the bench command,
# per-op cycle profiles, Fig-5c entry point on the shipped checkpoint/channels in both precision modes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench (default flags)"; timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-400
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --levels 16 --no-cpu-baseline --no-alt > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full (ALD kernel, 16 levels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 1 -c 1 -o gpurun_out/prof_ald -f python bench.py --steps 1 --warmup 1 --levels 16 --no-cpu-baseline --no-alt > gpurun_out/ncu_ald.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_ald.log
for p in tf32x3 tf32; do timeout 300 python tools/profile_ops.py 148 $p > gpurun_out/ops_profile_$p.txt 2>&1; sed -n 1,2p gpurun_out/ops_profile_$p.txt; done
for p in tf32x3 tf32; do
  echo "== Fig-5c full run, shipped checkpoint + shipped 100 CDL-C channels, $p"
  ( time timeout 1200 python -m score_based_channels_b200.test_score --ckpt fixtures_local/score-deepest-cdl-c.pt --out_dir gpurun_out/fig5c_$p --seed 1234 --no_plot --precision $p ) > gpurun_out/fig5c_$p.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/fig5c_$p.log
  python - <<PY
import sys, torch, numpy as np, os
sys.path.insert(0, '.')
from score_based_channels_b200 import dotmap_shim
dotmap_shim.install()
r = torch.load('gpurun_out/fig5c_$p/results.pt', weights_only=False)
np.savez_compressed('gpurun_out/fig5c_${p}_summary.npz', avg_nmse=r['avg_nmse'], best_nmse=r['best_nmse'], snr_range=r['snr_range'], final_per_channel=r['nmse_log'][0,0,:,-1,:])
os.remove('gpurun_out/fig5c_$p/results.pt')
PY
done
ls -la gpurun_out | head -40
