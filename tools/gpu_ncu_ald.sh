#!/bin/bash
# ncu --set full of the production ALD kernel inside a short bench run (16 levels)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 1 -c 1 -o gpurun_out/prof_ald -f python bench.py --steps 1 --warmup 1 --levels 16 --no-cpu-baseline --precision ${PREC:-tf32x3} > gpurun_out/ncu_ald.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_ald.log
