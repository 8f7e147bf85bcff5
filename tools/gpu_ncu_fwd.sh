#!/bin/bash
# ncu --set full capture (with source correlation) of one fused forward launch, B=148, per precision mode.
mkdir -p gpurun_out
for p in ${PRECS:-tf32x3 tf32}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 3 -c 1 -o gpurun_out/prof_fwd_$p -f python tools/profile_ops.py 148 $p > gpurun_out/ncu_full_$p.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full_$p.log
done
ls -la gpurun_out
