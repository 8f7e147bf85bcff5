#!/bin/bash
# A/B of experiment builds of engine 1 (SBC_LIB = alternative library, e.g. built with -DSBC_NTHREADS=512 -DSBC_MINCTAS=2
# -DSBC_MAXNS=2 into build/): forward time of a 296-sample launch, per-CTA cycle totals, and a short ALD run at B=256.
# Usage: gpurun -- 'bash tools/gpu_variants.sh [lib.so ...]'   ("" = the in-tree library)
mkdir -p gpurun_out
for lib in "" "$@"; do
  echo "== lib=${lib:-default}"
  SBC_LIB=${lib:+$PWD/$lib} timeout 300 python tools/e2_check.py 2>&1 | grep "tf32x3:"
  SBC_LIB=${lib:+$PWD/$lib} timeout 300 python tools/profile_ops.py 296 tf32x3 2>&1 | head -3
done
