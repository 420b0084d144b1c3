#!/bin/bash
# chain kernel timeline probe (tools/chain_probe.cu, built here): per-tile stamps of CTA 0
mkdir -p gpurun_out
P=tools/chain_probe
{
for v in 0 2; do for ex in 0 1; do for halo in 1 0; do
  echo "=== variant $v dbg_exec $ex halo $halo"
  timeout -s KILL 60 $P 9 224 $v $ex $halo
done; done; done
echo "=== variant 1 (inter-stack) halo 1"
timeout -s KILL 60 $P 9 224 1 0 1
echo "=== NS / NM knobs, variant 0 halo 1"
for ns in 2 3 4; do echo "-- NS=$ns"; DF3D_CHAIN_NS=$ns timeout -s KILL 60 $P 9 224 0 0 1 | grep run; done
for nm in 2 3; do echo "-- NM=$nm"; DF3D_CHAIN_NM=$nm timeout -s KILL 60 $P 9 224 0 0 1 | grep run; done
} > gpurun_out/probe.txt 2>&1
cat gpurun_out/probe.txt
