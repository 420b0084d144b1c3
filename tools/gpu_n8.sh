#!/bin/bash
# 8-GPU session: configs[2] weak scaling and configs[3] (100 000 frames strong-scaled) through torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_bench_c2.log 2> gpurun_out/n8_bench_c2.err; echo "config2 N=8 exit=$?"; tail -1 gpurun_out/n8_bench_c2.log | cut -c1-400
timeout -s KILL 600 $TR --master-port 29522 bench.py --gpus 8 --config 4 --steps 2 --warmup 3 > gpurun_out/n8_bench_c4.log 2> gpurun_out/n8_bench_c4.err; echo "config4 N=8 exit=$?"; tail -1 gpurun_out/n8_bench_c4.log | cut -c1-600; tail -3 gpurun_out/n8_bench_c4.err
