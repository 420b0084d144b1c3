#!/bin/bash
# the ncu half of tools/gpu_round.sh alone (gpurun copies back at most 64 MiB: fewer launches per full capture)
mkdir -p gpurun_out
timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --set full --clock-control none --import-source on \
    -k regex:conv_ -s 8 -c 10 -f -o gpurun_out/prof_conv \
    python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full conv exit=$?"
timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --set full --clock-control none \
    -k regex:'argmax|pack_points|triangulate|ba_gradient|ba_schur|ba_solve|ba_lsmr|ba_backsub|ba_step|proc_median|proc_apply' -c 12 -f -o gpurun_out/prof_tail \
    python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_tail.log 2>&1
echo "ncu full tail exit=$?"
timeout -s KILL 400 ncu --nvtx --nvtx-include "df3d_step/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit=$? lines=$(wc -l < gpurun_out/traffic.csv)"
ls -la gpurun_out/ | head -20; du -sh gpurun_out
