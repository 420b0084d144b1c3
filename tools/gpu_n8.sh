#!/bin/bash
# 8-GPU session: N = 1 on the same box, configs[2] weak scaling at N = 8 (as benched, and with the replicated bundle
# adjustment capped to 1 000 frames: what the growth of the BA problem with N costs), configs[3] (100 000 frames strong-scaled)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "scaling")}, "e2e", round(d["e2e"]["value"], 1), d["clocks"], d.get("bundle_adjust"))
except Exception as e:
    print("parse failed", e)
PY
}
timeout -s KILL 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/n8_bench_n1.log 2> gpurun_out/n8_bench_n1.err; echo "config2 N=1 exit=$?"; show gpurun_out/n8_bench_n1.log
timeout -s KILL 400 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_bench_c2.log 2> gpurun_out/n8_bench_c2.err; echo "config2 N=8 exit=$?"; show gpurun_out/n8_bench_c2.log
timeout -s KILL 600 $TR --master-port 29522 bench.py --gpus 8 --config 4 --steps 2 --warmup 3 > gpurun_out/n8_bench_c4.log 2> gpurun_out/n8_bench_c4.err; echo "config4 N=8 exit=$?"; show gpurun_out/n8_bench_c4.log; tail -3 gpurun_out/n8_bench_c4.err
