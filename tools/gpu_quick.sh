#!/bin/bash
# quick GPU iteration: hourglass parity tests, op table, short bench.  usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-quick}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then sel=(-k "$kexpr"); else sel=(); fi
timeout -s KILL 500 python -m pytest tests/test_gpu_hourglass.py tests/test_gpu_fullsize.py tests/test_gpu_conv.py -q -m gpu --no-header -rA -s "${sel[@]}" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|FAILED|ERROR|heat err|vs fp32" gpurun_out/${tag}_pytest.log | tail -30
timeout -s KILL 200 python tools/op_table.py > gpurun_out/${tag}_op_table.txt 2>&1; head -24 gpurun_out/${tag}_op_table.txt
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench exit=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'share', round(d['roofline']['share_of_step'],3), d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${tag}_bench.err').read()[-1500:])
PY
