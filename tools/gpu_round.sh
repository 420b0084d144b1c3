#!/bin/bash
# GPU session: full gpu test suite, smoke, bench, and (with "ncu") the ncu launch list of one bench
# step + full captures of the dominant kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu --no-header -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
( timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" ) | tee -a gpurun_out/summary.txt
( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?" ) | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench.log
( timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit=$?" ) | tee -a gpurun_out/summary.txt
tail -1 gpurun_out/bench_ref.log
if [ "$1" = "ncu" ]; then
  # launch list of ONE full-size bench step (NVTX range df3d_step; warm-up launches excluded)
  timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches exit=$? lines=$(wc -l < gpurun_out/launches.csv)" | tee -a gpurun_out/summary.txt
  # full captures (small batch: ncu replays each launch ~40x): a few conv launches of each shape class + the 2D->3D tail
  timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --set full --clock-control none --import-source on \
      -k regex:conv_ -s 8 -c 14 -f -o gpurun_out/prof_conv \
      python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full conv exit=$?" | tee -a gpurun_out/summary.txt
  timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --set full --clock-control none --import-source on \
      -k regex:'argmax|pack_points|triangulate|ba_gradient|ba_schur|ba_solve|ba_lsmr|ba_backsub|ba_step|proc_median|proc_apply' -c 16 -f -o gpurun_out/prof_tail \
      python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_tail.log 2>&1
  echo "ncu full tail exit=$?" | tee -a gpurun_out/summary.txt
fi
if [ "$1" = "traffic" ] || [ "$2" = "traffic" ]; then
  # DRAM bytes of every launch of one step (metrics pass, no --set full): roofline.traffic of bench.py
  timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_traffic.log 2>&1
  echo "ncu traffic exit=$? lines=$(wc -l < gpurun_out/traffic.csv)" | tee -a gpurun_out/summary.txt
fi
