#!/bin/bash
# per-kernel durations of the exact bundle adjustment at 8 000 frames (ncu launch list, one solve)
mkdir -p gpurun_out
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'ba_|triangulate' -c 120 --log-file gpurun_out/ba_launches.csv python tools/ba_time.py 8000 > gpurun_out/ba_ncu.log 2>&1
echo "exit=$? lines=$(wc -l < gpurun_out/ba_launches.csv)"
python tools/launch_summary.py gpurun_out/ba_launches.csv ba8000 | head -20
