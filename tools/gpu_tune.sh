#!/bin/bash
# BASELINE config 4 on one GPU (reduced to 25 channels): tune_hparams_score alpha x beta grid, full schedule
mkdir -p gpurun_out
( time timeout 1500 python -m score_based_channels_b200.tune_hparams_score --ckpt fixtures_local/score-deepest-cdl-c.pt --out_dir gpurun_out/tune --num_channels 25 --seed 7 --no_plot ) > gpurun_out/tune.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/tune.log
python - <<'PY'
import sys, torch, numpy as np, os
sys.path.insert(0, '.')
from score_based_channels_b200 import dotmap_shim
dotmap_shim.install()
r = torch.load('gpurun_out/tune/CDL-C-hyperparameters.pt', weights_only=False)
print('nmse_log', r['nmse_log'].shape, 'best_alpha_snr', r['best_alpha_snr'], 'best_beta_snr', r['best_beta_snr'])
print('best NMSE dB per SNR', np.round(10*np.log10(np.min(r['best_nmse'].reshape(-1, r['best_nmse'].shape[-1]), axis=0)), 2))
os.remove('gpurun_out/tune/CDL-C-hyperparameters.pt')
PY
