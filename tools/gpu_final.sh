#!/bin/bash
# last session of the round: whole gpu suite, smoke, configs[2] bench + reference arm, BA timing
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests -q -m gpu --no-header -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -5
( timeout -s KILL 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" ) | tee -a gpurun_out/summary.txt
( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?" ) | tee -a gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-300
( timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit=$?" ) | tee -a gpurun_out/summary.txt
timeout -s KILL 100 python tools/ba_time.py 256 1000 8000 2>&1 | grep "^T=" | tee gpurun_out/ba_time.txt
timeout -s KILL 100 python tools/ba_sharded_time.py 8000 2>&1 | grep "^T=" | tee gpurun_out/ba_sharded_time_1gpu.txt
