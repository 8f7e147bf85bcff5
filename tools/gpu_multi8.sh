#!/bin/bash
# N-GPU validation (N = number of visible GPUs): config 2 weak scaling and config 4 strong scaling on truncated schedules
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 2 --warmup 3 --levels 48 --no-extra > gpurun_out/bench_n${N}_cfg2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n${N}_cfg2.log | cut -c1-240
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config 4 --engine 1 --steps 2 --warmup 3 --levels 6 --no-extra > gpurun_out/bench_n${N}_cfg4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n${N}_cfg4.log | cut -c1-240
