"""Device-side composition of the hot path: images -> hourglass -> arg-max -> 19->38 packing ->
DLT -> bundle adjustment -> DLT, everything enqueued on one CUDA stream with no host sync.

This is what ``Core.pose2d_estimation`` + ``Core.calibrate_calc`` + ``Core.save`` run in the
reference (df3d/core.py:170-203, 229-250, 351-360) minus file I/O, and what bench.py times.
Frames shard across ranks (``run_sharded``): every stage is per-frame except bundle adjustment,
whose Schur-reduced camera system is all-reduced each iteration, and the final all-gather of the
3-D joints.
"""
import numpy as np
import torch

from . import ops
from .camera_network import rodrigues_vec
from .hourglass import HourglassEngine
from .ops import intr_to_vec4
from .skeleton import NUM_CAMERAS


def load_default_calib():
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "calib.npz")
    with np.load(path) as z:
        return {k: z[k].astype(np.float64) for k in ("R", "tvec", "intr", "distort")}


def reorder_calib(calib, camera_ordering):
    """calib_reordered[cidx] = calib[idx] for (idx, cidx) in enumerate(camera_ordering) (core.py:240-242)."""
    out = {k: np.empty_like(v) for k, v in calib.items()}
    for idx, cidx in enumerate(camera_ordering):
        for k in out:
            out[k][cidx] = calib[k][idx]
    return out


class Pose3DPipeline:
    def __init__(self, state_dict, in_h, in_w, max_images, image_shape, camera_ordering=range(7), calib=None,
                 device="cuda", mean=0.5, ba_max_iters=10, ba_ftol=1e-4):
        self.device = torch.device(device)
        self.engine = HourglassEngine(state_dict, in_h, in_w, max_images, device=device, mean=mean)
        self.order = [int(c) for c in camera_ordering]
        self.image_shape = [int(image_shape[0]), int(image_shape[1])]  # [W, H] like Core.image_shape
        calib = reorder_calib(load_default_calib() if calib is None else calib, self.order)
        self.calib = calib
        cam_rt = np.stack([np.concatenate([rodrigues_vec(calib["R"][c]), calib["tvec"][c]]) for c in range(NUM_CAMERAS)])
        self.cam_rt0 = torch.as_tensor(cam_rt, device=self.device)
        self.intr4 = torch.as_tensor(intr_to_vec4(calib["intr"]), device=self.device)
        self.ba_max_iters, self.ba_ftol = ba_max_iters, ba_ftol
        self._flip_cache = {}
        self._ba_ws = {}

    def flip_flags(self, T):
        """Images are camera-major (b = c*T + t); cameras camera_ordering[4:] are mirrored (core.py:179)."""
        if T not in self._flip_cache:
            f = torch.zeros((NUM_CAMERAS, T), dtype=torch.uint8)
            for c in self.order[4:]:
                f[c] = 1
            self._flip_cache[T] = f.reshape(-1).to(self.device)
        return self._flip_cache[T]

    def pose2d(self, images, T):
        """images (7*T,H,W) uint8 CUDA, camera-major -> idx (7*T,K), conf (7*T,K), points2d, pts_xy."""
        idx, conf = self.engine.forward(images, flip=self.flip_flags(T))
        p2d, pxy = ops.pack_points2d(idx, NUM_CAMERAS, T, self.engine.heatmap_shape, self.order, self.image_shape)
        return idx, conf, p2d, pxy

    def pose3d(self, pxy, group=None):
        """pts_xy (7,T,J,2) -> cameras after BA (7,6), points3d (T,J,3) re-triangulated (core.py:355)."""
        Cn, T, J, _ = pxy.shape
        cam = self.cam_rt0.clone()
        P0, _ = ops.projection_matrices(cam, self.intr4)
        X = ops.triangulate_dlt(P0, pxy)
        if group is not None:
            rep = ops.bundle_adjust_distributed(cam, self.intr4, pxy, X, group=group, max_iters=self.ba_max_iters, ftol=self.ba_ftol)
        else:
            key = (Cn, T, J)
            if key not in self._ba_ws:
                self._ba_ws[key] = ops.ba_workspace(Cn, T, J, self.device)
            rep = ops.bundle_adjust(cam, self.intr4, pxy, X, max_iters=self.ba_max_iters, ftol=self.ba_ftol,
                                    workspace=self._ba_ws[key])
        P1, R1 = ops.projection_matrices(cam, self.intr4)
        X1 = ops.triangulate_dlt(P1, pxy)
        return cam, R1, X1, rep

    def run(self, images, T, group=None):
        idx, conf, p2d, pxy = self.pose2d(images, T)
        cam, R1, X1, rep = self.pose3d(pxy, group=group)
        return {"idx": idx, "conf": conf, "points2d": p2d, "pts_xy": pxy, "cam_rt": cam, "R": R1,
                "points3d_wo_procrustes": X1, "ba_report": rep}

    def launches(self, n_images):
        """Kernel launches of one `run` (for bench.py's gpu_launches): hourglass plan + pack + 2 x
        (projection + DLT) + BA (begin + max_iters x 5 + end)."""
        return self.engine.launches(n_images) + 1 + 4 + (2 + 5 * self.ba_max_iters)


def gather_frames(x, group=None):
    """Single all-gather of a per-rank (T_local, ...) tensor along the frame axis (NCCL over NVLink
    on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=0)


def shard_frames(T, rank, world):
    """Contiguous frame block of `rank` (all 7 cameras of a frame stay on one rank)."""
    per = (T + world - 1) // world
    lo = min(T, rank * per)
    return lo, min(T, lo + per)
