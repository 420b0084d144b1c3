#!/bin/bash
# chain timeline probes (tools/chain_probe.cu): T1 chain with / without the up-sampled residual, inter-stack form
mkdir -p gpurun_out
P=tools/chain_probe
{
echo "=== T1 +res+up";  timeout -s KILL 60 $P 9 448 0 0 1 0
echo "=== T1 +res";     timeout -s KILL 60 $P 9 448 0 0 1 2
echo "=== T3 (5-stage inter-stack form)"; timeout -s KILL 60 $P 9 448 1 0 1 0
echo "=== T1 +res+up, accumulator-complete stamps"; timeout -s KILL 60 $P 9 448 0 1 1 0
} > gpurun_out/probe4.txt 2>&1
cat gpurun_out/probe4.txt
