#!/bin/bash
# chain-kernel iteration: timeline probe first (short, catches hangs), then the hourglass parity tests, op table, short bench
tag=${1:-q2}
mkdir -p gpurun_out
{ echo "=== T1 +res+up"; timeout -s KILL 60 tools/chain_probe 9 448 0 4 1 0; echo "exit=$?"; } > gpurun_out/${tag}_probe.txt 2>&1
grep -E "run 2|exit=|tile 2[12]:|slab 8" gpurun_out/${tag}_probe.txt | head -12
if grep -q "exit=0" gpurun_out/${tag}_probe.txt; then bash tools/gpu_quick.sh $tag "$2"; fi
