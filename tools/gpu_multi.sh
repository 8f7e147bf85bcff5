#!/bin/bash
# 2-GPU validation: (1) the entry point sharded over 2 ranks gives the same results.pt as 1 rank (same seed),
# (2) bench.py under torchrun prints its line (weak scaling, NCCL all-gather of the NMSE logs).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
CK=fixtures_local/score-deepest-cdl-c.pt
timeout 600 python -m score_based_channels_b200.test_score --ckpt $CK --out_dir gpurun_out/ts1 --levels 4 --num_channels 10 --seed 5 --no_plot > gpurun_out/ts1.log 2>&1; echo "rc1=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 -m score_based_channels_b200.test_score --ckpt $CK --out_dir gpurun_out/ts2 --levels 4 --num_channels 10 --seed 5 --no_plot > gpurun_out/ts2.log 2>&1; echo "rc2=$?"; tail -3 gpurun_out/ts2.log
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from score_based_channels_b200 import dotmap_shim
dotmap_shim.install()
a=torch.load('gpurun_out/ts1/results.pt',weights_only=False); b=torch.load('gpurun_out/ts2/results.pt',weights_only=False)
print('shard invariance: nmse_log identical =', np.array_equal(a['nmse_log'],b['nmse_log']), 'max abs diff', float(np.abs(a['nmse_log']-b['nmse_log']).max()), a['nmse_log'].shape)
PY
rm -rf gpurun_out/ts1/results.pt gpurun_out/ts2/results.pt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --levels 48 --no-extra > gpurun_out/bench_n2.log 2>&1; echo "rc3=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-700
