#!/bin/bash
mkdir -p gpurun_out
P=tools/chain_probe
{
echo "=== baseline"; timeout -s KILL 30 $P 9 224 0 0 1
echo "=== no stores"; timeout -s KILL 30 $P 9 224 0 2 1
echo "=== no up-res"; timeout -s KILL 30 $P 9 224 0 0 1 2
echo "=== no stores, no up-res"; timeout -s KILL 30 $P 9 224 0 2 1 2
echo "=== small batch 16 (L2 resident)"; timeout -s KILL 30 $P 9 16 0 0 1
} > gpurun_out/probe2.txt 2>&1
grep -E "===|run 2|tile 2[12]:" gpurun_out/probe2.txt
