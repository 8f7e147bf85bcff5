#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2_check.py 2>&1 | grep "tf32x3:"
echo "instrumented build (SBC_DBG=32):"; SBC_DBG=32 timeout 300 python tools/e2_check.py 2>&1 | grep "tf32x3:"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbc_ald_kernel -s 1 -c 1 -o gpurun_out/prof_e1x2 -f python tools/e2_short.py 296 2 tf32x3 > gpurun_out/ncu_e1x2.log 2>&1; echo "ncu rc=$?"
