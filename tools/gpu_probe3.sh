#!/bin/bash
mkdir -p gpurun_out
P=tools/chain_probe
{
for cfg in "3 5" "3 4" "3 3" "4 3" "2 6" "2 5"; do set -- $cfg
echo "=== NS=$1 NM=$2"; DF3D_CHAIN_NS=$1 DF3D_CHAIN_NM=$2 timeout -s KILL 30 $P 9 224 0 0 1
done
} > gpurun_out/probe3.txt 2>&1
grep -E "===|run 2|tile 2[12]:" gpurun_out/probe3.txt
