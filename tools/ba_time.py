"""Device time of bundle adjustment + DLT + procrustes on the config-3 geometry at several frame counts (GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfly3d_b200 import ops
from deepfly3d_b200.ops import intr_to_vec4
from oracle import geometry as g, synth

for T in [int(a) for a in sys.argv[1:]] or [256, 1000, 8000]:
    calib, pts, _ = synth.config3_geometry(T, seed=2)
    cam0 = np.stack([np.concatenate([g.rodrigues_inv(calib["R"][k]), calib["tvec"][k]]) for k in range(7)])
    intr4 = torch.as_tensor(intr_to_vec4(calib["intr"])).cuda()
    pxy = torch.as_tensor(pts).cuda()
    ws = ops.ba_workspace(7, T, 38, pxy.device)
    for solver in ("lsmr", "exact"):
        max_iters = 10
        ts = []
        for rep in range(4):
            cam = torch.as_tensor(cam0).cuda()
            P0, _ = ops.projection_matrices(cam, intr4)
            X = ops.triangulate_dlt(P0, pxy)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rep_ = ops.bundle_adjust(cam, intr4, pxy, X, max_iters=max_iters, workspace=ws, solver=solver)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        r = ops.ba_report(rep_)
        e0.record(); X1 = ops.triangulate_dlt(P0, pxy); e1.record(); torch.cuda.synchronize(); t_dlt = e0.elapsed_time(e1)
        e0.record(); ops.procrustes(X1); e1.record(); torch.cuda.synchronize(); t_pr = e0.elapsed_time(e1)
        print(f"T={T} {solver}: BA {min(ts):.3f} ms ({r['iters']} evaluations, status {r['status']}, last lsmr itn {r['lsmr_itn']}), DLT {t_dlt:.3f} ms, procrustes {t_pr:.3f} ms")
