#!/bin/bash
# timing experiments: which part of an op costs what (SBC_DBG bits, results are garbage)
for d in 0 1 2 3 4 7 16; do echo "== SBC_DBG=$d"; SBC_DBG=$d timeout 120 python tools/profile_ops.py 148 tf32x3 > gpurun_out/dbg_$d.txt 2>&1; sed -n 2p gpurun_out/dbg_$d.txt; grep -E "^ +(5|9|16|18|36|57|58|60|61) " gpurun_out/dbg_$d.txt | cut -c1-150; done
