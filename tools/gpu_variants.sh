#!/bin/bash
mkdir -p gpurun_out
for d in 0 150000 300000 550000; do
  echo "== SBC_DESYNC=$d"; SBC_DESYNC=$d timeout 300 python bench.py --levels 48 --steps 2 --warmup 1 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.2f est/s ms/step %.2f'%(d['value'], d['ms_per_step']))"
done
echo "== config 5, engine 1 (global arena, 2 CTAs/SM)"; timeout 300 python bench.py --config 5 --engine 1 --batch 2368 --levels 3 --steps 2 --warmup 1 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.2f est/s ms/step %.2f ctas %s'%(d['value'], d['ms_per_step'], d['roofline']['ctas_per_sm']))"
