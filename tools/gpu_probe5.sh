#!/bin/bash
mkdir -p gpurun_out
P=tools/chain_probe
{
echo "=== T1 +res+up baseline";  timeout -s KILL 60 $P 9 448 0 0 1 0
echo "=== T1 +res+up, stores skipped (dbg_exec 2)";  timeout -s KILL 60 $P 9 448 0 2 1 0
echo "=== T1, no residual slabs (nores 1)";  timeout -s KILL 60 $P 9 448 0 0 1 1
echo "=== T1 1x1 head (taps 1)";  timeout -s KILL 60 $P 1 448 0 0 1 0
} > gpurun_out/probe5.txt 2>&1
grep -E "===|run 2|tile 2[12]:" gpurun_out/probe5.txt
