#!/bin/bash
# bundle-adjustment iteration: geometry tests (incl. the sharded solver simulated on one GPU); with 2 GPUs also the NCCL test
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_fullsize.py -q -m gpu --no-header -rA -s -k "bundle or sharded or calib or golden or reproducible" > gpurun_out/ba_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|FAILED|ERROR|differ|Error" gpurun_out/ba_pytest.log | tail -15
if [ "$(python -c 'import torch; print(torch.cuda.device_count())')" -ge 2 ]; then
  timeout -s KILL 600 python -m pytest tests/test_gpu_core.py -q -m gpu --no-header -rA -s -k "sharded_two" > gpurun_out/ba_pytest_n2.log 2>&1; echo "pytest n2 exit=$?"
  grep -E "passed|failed|FAILED|ERROR|differ|Error" gpurun_out/ba_pytest_n2.log | tail -15
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/ba_sharded_time.py 8000 2>&1 | grep "^T=" | tee gpurun_out/ba_sharded_time.txt
fi
timeout -s KILL 100 python tools/ba_sharded_time.py 8000 2>&1 | grep "^T=" | tee -a gpurun_out/ba_sharded_time.txt
timeout -s KILL 100 python tools/ba_time.py 256 1000 8000 2>&1 | grep "^T=" | tee gpurun_out/ba_time.txt
