#!/bin/bash
# images per launch sequence (DF3D_HG_CHUNK) against the step time of configs[2] (7 000 images per step)
mkdir -p gpurun_out
for c in 1792 2334 3500 1792; do
  DF3D_HG_CHUNK=$c timeout -s KILL 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/chunk_$c.log 2> gpurun_out/chunk_$c.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/chunk_$c.log').read().strip().splitlines()[-1])
    print('chunk $c', round(d['value'],1), round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e:
    print('chunk $c failed', e); print(open('gpurun_out/chunk_$c.err').read()[-800:])
PY
done
