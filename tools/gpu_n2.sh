#!/bin/bash
# 2-GPU session: the whole gpu test suite (incl. the NCCL sharded == single-GPU test) and the N = 2 bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --no-header -rA -s > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|SKIPPED|FAILED|ERROR" gpurun_out/pytest_gpu_n2.log | tail -8
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?"
tail -1 gpurun_out/bench_n2.log | cut -c1-400
