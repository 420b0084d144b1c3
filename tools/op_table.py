"""Per-launch table of the hourglass plan on the GPU box: device ms, TFLOP/s, GB/s per op class.
    python tools/op_table.py [frames] > gpurun_out/op_table.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfly3d_b200.hourglass import HourglassEngine
from deepfly3d_b200.inference import random_state_dict

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = 7 * frames
eng = HourglassEngine(random_state_dict(8, seed=0), 256, 256, n)
img = torch.randint(0, 255, (n, 256, 256), dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.forward(img)
torch.cuda.synchronize()
eng.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.forward(img)
e1.record()
torch.cuda.synchronize()
rows = eng.op_table()
total = sum(r[1] for r in rows)
print(f"images {n}  forward {e0.elapsed_time(e1):.2f} ms  sum of ops {total:.2f} ms  ops/chunk {len(rows)}")
agg = collections.OrderedDict()
for label, ms, fl, by in rows:
    a = agg.setdefault(label, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += fl; a[3] += by
print(f"{'op class':46s} {'n':>4s} {'ms':>9s} {'%':>6s} {'TFLOP/s':>9s} {'GB/s':>8s} {'us/launch':>10s}")
chunks = (n + 127) // 128
for label, (cnt, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{label:46s} {cnt:4d} {ms:9.3f} {100*ms/total:6.2f} {fl/ms/1e9 if ms else 0:9.1f} {by/ms/1e6 if ms else 0:8.0f} {1e3*ms/cnt/chunks:10.1f}")
