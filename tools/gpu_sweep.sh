#!/bin/bash
# ring-depth sweeps (profiling knobs) -> gpurun_out/sweep.txt
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
for st in 2 3 4 5; do
  echo "== unfused, DF3D_CONV_STAGES=$st" >> gpurun_out/sweep.txt
  DF3D_HG_FUSE=0 DF3D_CONV_STAGES=$st timeout -s KILL 200 python tools/op_table.py 128 2>&1 | grep -E "forward|conv3x3   64x64|conv1x1   64x64   cin=256 BN=128" >> gpurun_out/sweep.txt
done
for cfg in "8 2 2" "7 2 3" "6 2 4" "6 4 2" "5 4 3" "4 4 4" "4 3 4" "3 4 4"; do
  set -- $cfg
  echo "== fuse 2, NM=$1 NW=$2 NS=$3" >> gpurun_out/sweep.txt
  DF3D_HG_FUSE=2 DF3D_CHAIN_NM=$1 DF3D_CHAIN_NW=$2 DF3D_CHAIN_NS=$3 timeout -s KILL 200 python tools/op_table.py 128 2>&1 | grep -E "forward|chain3x3   64x64" >> gpurun_out/sweep.txt
done
cat gpurun_out/sweep.txt
