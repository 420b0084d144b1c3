#!/bin/bash
# ingest iteration: JPEG / loader tests and the Core-from-files leg of the bench with block-size variants.  usage: tools/gpu_ingest.sh <tag>
tag=${1:-ingest}
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_core.py -q -m gpu --no-header -rA -s -k "jpeg or device_decode or streams_blocks" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|FAILED|ERROR|SKIPPED|nvJPEG|backend" gpurun_out/${tag}_pytest.log | tail -20
DF3D_BENCH_BLOCKS=${DF3D_BENCH_BLOCKS:-32,64} timeout -s KILL 600 python bench.py --steps 1 --warmup 3 --frames 256 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench exit=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
    for k,v in d['e2e_files']['variants'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${tag}_bench.err').read()[-1500:])
PY
