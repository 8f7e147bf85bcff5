#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (entry)"; timeout 900 python -m pytest tests/test_gpu_entry.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_entry.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_entry.log
echo "== Fig-5c full run, shipped checkpoint + shipped 100 CDL-C channels, tf32x3"
( time timeout 1200 python -m score_based_channels_b200.test_score --ckpt fixtures_local/score-deepest-cdl-c.pt --out_dir gpurun_out/fig5c_tf32x3 --seed 1234 --no_plot ) > gpurun_out/fig5c_tf32x3.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/fig5c_tf32x3.log
rm -f gpurun_out/fig5c_tf32x3/results.pt.keep; python - <<'PY'
import torch, numpy as np
r = torch.load('gpurun_out/fig5c_tf32x3/results.pt', weights_only=False)
np.savez_compressed('gpurun_out/fig5c_tf32x3_summary.npz', avg_nmse=r['avg_nmse'], best_nmse=r['best_nmse'], snr_range=r['snr_range'],
                    final_per_channel=r['nmse_log'][0,0,:,-1,:])
import os; os.remove('gpurun_out/fig5c_tf32x3/results.pt')
PY
