#!/bin/bash
# bring-up of the CTA-pair chain kernel: probe first (short timeout), then parity tests
mkdir -p gpurun_out
{
echo "=== probe variant 0 halo 1"; timeout -s KILL 30 tools/chain_probe 9 224 0 0 1; echo "rc=$?"
echo "=== probe variant 0 halo 0"; timeout -s KILL 30 tools/chain_probe 9 224 0 0 0; echo "rc=$?"
echo "=== probe variant 1 halo 1"; timeout -s KILL 30 tools/chain_probe 9 224 1 0 1; echo "rc=$?"
} > gpurun_out/probe_pair.txt 2>&1
cat gpurun_out/probe_pair.txt
timeout -s KILL 150 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hourglass.py -q -m gpu -x --no-header > gpurun_out/pytest_pair.log 2>&1; echo "pytest exit=$?"
tail -15 gpurun_out/pytest_pair.log
