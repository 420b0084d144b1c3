"""Device time of the exact bundle adjustment over T frames: replicated on every rank against frame-sharded
(ops.bundle_adjust_sharded).  torchrun --nproc-per-node N tools/ba_sharded_time.py [T]   (noisy points: all 10 evaluations run)"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfly3d_b200 import ops
from deepfly3d_b200.ops import intr_to_vec4
from oracle import geometry as g, synth

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
calib, pts, _ = synth.config3_geometry(T, seed=2)
rng = np.random.default_rng(0)
pts = np.where(pts != 0, pts + rng.normal(0, 30.0, pts.shape), pts)          # garbage-like observations: no convergence in 10 evaluations
cam0 = np.stack([np.concatenate([g.rodrigues_inv(calib["R"][k]), calib["tvec"][k]]) for k in range(7)])
intr4 = torch.as_tensor(intr_to_vec4(calib["intr"])).cuda()
pxy = torch.as_tensor(pts).cuda()
ws = ops.ba_workspace(7, T, 38, pxy.device)
group = dist.group.WORLD if world > 1 else None
for mode in ("replicated", "sharded"):
    ts = []
    for rep in range(5):
        cam = torch.as_tensor(cam0).cuda()
        P0, _ = ops.projection_matrices(cam, intr4)
        X = ops.triangulate_dlt(P0, pxy)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "sharded":
            r_ = ops.bundle_adjust_sharded(cam, intr4, pxy, X, group=group, max_iters=10, workspace=ws)
        else:
            r_ = ops.bundle_adjust(cam, intr4, pxy, X, max_iters=10, workspace=ws, solver="exact")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    r = ops.ba_report(r_)
    if rank == 0:
        print(f"T={T} world={world} {mode}: BA {min(ts):.3f} ms (all runs {[round(t, 2) for t in ts]}; {r['iters']} evaluations, status {r['status']})", flush=True)
if world > 1:
    dist.destroy_process_group()
