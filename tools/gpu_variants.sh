#!/bin/bash
# chain-length experiment: forward time at B=148 (one sample per SM) for experiment builds of engine 1
mkdir -p gpurun_out
for lib in "" build/lib_512_64.so build/lib_256.so; do
  echo "== lib=${lib:-default}"
  SBC_LIB=${lib:+$PWD/$lib} timeout 300 python tools/profile_ops.py 148 tf32x3 2>&1 | head -2
done
echo "== engine 2 check"; timeout 300 python tools/e2_check.py 2>&1 | tail -4
