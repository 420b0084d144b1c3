mkdir -p gpurun_out
timeout -s KILL 600 ncu --nvtx --nvtx-include "df3d_step/" --set full --clock-control none --import-source on -k regex:conv_chain -s 4 -c 14 -f -o gpurun_out/r02h_prof_chain python bench.py --profile --frames 256 --steps 1 --warmup 3 > gpurun_out/r02h_ncu_chain.log 2>&1
echo "ncu chain exit=$?"; ls -la gpurun_out/r02h_prof_chain.ncu-rep
timeout -s KILL 400 python -m pytest tests/test_gpu_hourglass.py -q -m gpu --no-header -rA -s -k "trained" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "trained 2-stack|passed|failed|Error|assert" gpurun_out/r02h_pytest.log | head
