#!/bin/bash
mkdir -p gpurun_out
P=tools/chain_probe
{
echo "=== T1 +res+up, slab stamps";  timeout -s KILL 60 $P 9 448 0 4 1 0
echo "=== inter-stack 5-stage form, slab stamps";  timeout -s KILL 60 $P 9 448 1 4 1 0
} > gpurun_out/probe6.txt 2>&1
grep -E "===|run 2|tile 2[12]:|slab" gpurun_out/probe6.txt
