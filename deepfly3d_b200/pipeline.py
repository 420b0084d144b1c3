"""Device-side composition of the hot path: images -> hourglass -> arg-max -> 19->38 packing ->
DLT -> bundle adjustment -> DLT, everything enqueued on one CUDA stream with no host sync.

This is what ``Core.pose2d_estimation`` + ``Core.calibrate_calc`` + ``Core.save`` run in the
reference (df3d/core.py:170-203, 229-250, 351-360) minus file I/O, and what bench.py times.

Multi-GPU (one process per GPU, ``group`` = the NCCL process group): frames shard in contiguous blocks, all
7 cameras of a frame on one rank.  Hourglass, arg-max, packing and DLT need no communication.  Bundle
adjustment is ONE global problem over all frames (42 shared camera unknowns): the packed 2-D
points of the frames it uses (4 256 B / frame) are all-gathered and every rank ends with identical cameras.  With
the LSMR solver every rank solves the whole problem with the same bit-reproducible kernels (no collective inside the
solver); with the exact solver the per-point work of each pass is split by blocks of points between the ranks and
the per-block partial sums are all-gathered and summed in the single-GPU order (ops.bundle_adjust_sharded) -- in
both cases the cameras equal the single-GPU ones to the bit.  Each rank then triangulates its own frames; one
all-gather of the 3-D joints (912 B / frame) ends the step.
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
                 device="cuda", mean=0.5, ba_max_iters=10, ba_ftol=1e-4, ba_max_frames=None, ba_solver="lsmr"):
        self.device = torch.device(device)
        self.engine = HourglassEngine(state_dict, in_h, in_w, max_images, device=device, mean=mean)
        self.order = [int(c) for c in camera_ordering]
        self.image_shape = [int(image_shape[0]), int(image_shape[1])]  # [W, H] like Core.image_shape
        calib = reorder_calib(load_default_calib() if calib is None else calib, self.order)
        self.calib = calib
        cam_rt = np.stack([np.concatenate([rodrigues_vec(calib["R"][c]), calib["tvec"][c]]) for c in range(NUM_CAMERAS)])
        self.cam_rt0 = torch.as_tensor(cam_rt, device=self.device)
        self.intr4 = torch.as_tensor(intr_to_vec4(calib["intr"]), device=self.device)
        self.ba_max_iters, self.ba_ftol, self.ba_max_frames, self.ba_solver = ba_max_iters, ba_ftol, ba_max_frames, ba_solver
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

    def ba_points(self, pxy, group=None):
        """The 2-D points the bundle adjustment runs on, identical on every rank: all frames of all ranks (what the
        reference does), or -- `ba_max_frames` set and exceeded -- every stride-th frame of the whole recording
        (100 000 frames, BA on <= 1 000).  Ranks select their own frames first, so only
        the subset crosses NVLink; a rank with one frame fewer pads with an unobserved frame (all zeros)."""
        import torch.distributed as dist

        T = pxy.shape[1]
        world = dist.get_world_size(group) if group is not None else 1
        rank = dist.get_rank(group) if group is not None else 0
        total = T * world
        if self.ba_max_frames is not None and total > self.ba_max_frames:
            stride = -(-total // self.ba_max_frames)
            sel = torch.arange((-rank * T) % stride, T, stride, device=pxy.device)
            n_max = -(-T // stride)
            local = torch.zeros((pxy.shape[0], n_max) + tuple(pxy.shape[2:]), dtype=pxy.dtype, device=pxy.device)
            local[:, :sel.numel()] = pxy.index_select(1, sel)
        else:
            local = pxy
        return gather_frames(local, group, dim=1) if group is not None else local.contiguous()

    def pose3d(self, pxy, group=None):
        """pts_xy (7,T,J,2) of this rank's frames -> cameras after BA (7,6), R (7,3,3), points3d (T,J,3)
        re-triangulated with the new cameras (core.py:355), BA report."""
        Cn, T, J, _ = pxy.shape
        cam = self.cam_rt0.clone()
        ba_xy = self.ba_points(pxy, group)
        P0, _ = ops.projection_matrices(cam, self.intr4)
        X = ops.triangulate_dlt(P0, ba_xy)
        key = (Cn, ba_xy.shape[1], J)
        if key not in self._ba_ws:
            self._ba_ws[key] = ops.ba_workspace(*key, self.device)
        world = 1
        if group is not None:
            import torch.distributed as dist

            world = dist.get_world_size(group)
        if world > 1 and self.ba_solver == "exact" and ops.ba_sharded_plan(*key, world) is not None:
            # the per-point work of every pass split between the ranks, per-block partial sums all-gathered and summed in
            # the single-GPU order: same cameras to the bit, 1 / world of the work (the LSMR solver runs replicated)
            rep = ops.bundle_adjust_sharded(cam, self.intr4, ba_xy, X, group=group, max_iters=self.ba_max_iters, ftol=self.ba_ftol,
                                            workspace=self._ba_ws[key])
        else:
            rep = ops.bundle_adjust(cam, self.intr4, ba_xy, X, max_iters=self.ba_max_iters, ftol=self.ba_ftol,
                                    workspace=self._ba_ws[key], solver=self.ba_solver)
        P1, R1 = ops.projection_matrices(cam, self.intr4)
        X1 = ops.triangulate_dlt(P1, pxy)                        # own frames only
        return cam, R1, X1, rep

    def run(self, images, T, group=None):
        idx, conf, p2d, pxy = self.pose2d(images, T)
        cam, R1, X1, rep = self.pose3d(pxy, group=group)
        return {"idx": idx, "conf": conf, "points2d": p2d, "pts_xy": pxy, "cam_rt": cam, "R": R1,
                "points3d_wo_procrustes": X1, "ba_report": rep}

    def launches(self, n_images, sharded_ba=False):
        """Kernel launches of one `run` (for bench.py's gpu_launches): hourglass plan + pack + 2 x
        (projection + DLT) + bundle adjustment (sharded exact solver: per iteration 4 passes + 4 finishes + solve + apply)."""
        ba = 2 + 10 * self.ba_max_iters if sharded_ba else ops.bundle_adjust_launches(self.ba_max_iters, self.ba_solver)
        return self.engine.launches(n_images) + 1 + 4 + ba


def gather_frames(x, group=None, dim=0):
    """Single all-gather of a per-rank tensor along its frame axis `dim` (NCCL over NVLink on the GPU box,
    gloo in the CPU tests).  Every rank must hold the same number of frames."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)          # concatenation along dim 0, rank-major
    if dim == 0:
        return out
    # (world, ..., T, ...) -> (..., world * T, ...): rank-major along the frame axis
    out = out.reshape((world,) + tuple(x.shape)).movedim(0, dim)
    shape = list(x.shape)
    shape[dim] *= world
    return out.reshape(shape).contiguous()


def shard_frames(T, rank, world):
    """Contiguous frame block of `rank` (all 7 cameras of a frame stay on one rank)."""
    per = (T + world - 1) // world
    lo = min(T, rank * per)
    return lo, min(T, lo + per)
