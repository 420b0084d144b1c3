#!/bin/bash
# GPU session: full gpu test suite, smoke, bench, ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu --no-header -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
( timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" ) | tee -a gpurun_out/summary.txt
( timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench.log
if [ "$1" = "ncu" ]; then
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 0 --frames 32 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches exit=$?" | tee -a gpurun_out/summary.txt
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 40 -c 6 -o gpurun_out/prof_conv \
      python bench.py --steps 1 --warmup 0 --frames 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?" | tee -a gpurun_out/summary.txt
fi
