#!/bin/bash
# quick timing: short bench for both precision modes + forward-only timing
mkdir -p gpurun_out
for p in tf32x3 tf32; do
  timeout 600 python bench.py --levels 24 --steps 2 --warmup 1 --no-cpu-baseline --precision $p > gpurun_out/bench_short_$p.log 2>&1; echo "rc=$?"
  python - <<PY
import json
l=open('gpurun_out/bench_short_$p.log').read().strip().split('\n')[-1]
d=json.loads(l); print('$p', 'value %.2f est/s  ms/step %.2f  e2e %.2f  frac %.4f clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks']))
PY
done
for p in tf32x3 tf32; do timeout 300 python tools/profile_ops.py 148 $p 2>&1 | head -2; done
