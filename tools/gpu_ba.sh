#!/bin/bash
# bundle-adjustment iteration: geometry tests (incl. the sharded solver simulated on one GPU); with 2 GPUs also the NCCL test
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_fullsize.py -q -m gpu --no-header -rA -s -k "bundle or sharded or calib or golden or reproducible" > gpurun_out/ba_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|FAILED|ERROR|differ|Error" gpurun_out/ba_pytest.log | tail -15
if [ "$(python -c 'import torch; print(torch.cuda.device_count())')" -ge 2 ]; then
  timeout -s KILL 600 python -m pytest tests/test_gpu_core.py -q -m gpu --no-header -rA -s -k "sharded_two" > gpurun_out/ba_pytest_n2.log 2>&1; echo "pytest n2 exit=$?"
  grep -E "passed|failed|FAILED|ERROR|differ|Error" gpurun_out/ba_pytest_n2.log | tail -15
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/ba_bench_n2.log 2> gpurun_out/ba_bench_n2.err; echo "bench n2 exit=$?"
  tail -1 gpurun_out/ba_bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['bundle_adjust'], d['gpu_launches'])" || tail -5 gpurun_out/ba_bench_n2.err
fi
