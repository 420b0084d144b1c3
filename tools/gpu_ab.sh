#!/bin/bash
# same-box A/B of two builds of the library (deepfly3d_b200/libA.so.bin, libB.so.bin): bench each twice, interleaved
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for r in 1 2; do for v in A B; do
  cp deepfly3d_b200/lib$v.so.bin deepfly3d_b200/libdf3d_b200.so
  timeout -s KILL 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'fps', d['clocks']['sm_mhz'], 'MHz')" >> gpurun_out/ab.txt
done; done
cp deepfly3d_b200/libB.so.bin deepfly3d_b200/libdf3d_b200.so
cat gpurun_out/ab.txt
