"""Thin Python wrappers: PyTorch owns device memory and streams, the C ABI does the work.

Every function takes CUDA tensors, enqueues on the current torch stream and returns CUDA tensors
without synchronising.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("deepfly3d_b200 ops take CUDA tensors (there is no CPU fallback)")


def heatmap_argmax(hm):
    """(B,K,H,W) float32/bfloat16 CUDA -> idx (B,K) int32, conf (B,K) float32."""
    _need_cuda(hm)
    if hm.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("heatmap_argmax: dtype must be float32 or bfloat16")
    hm = hm.contiguous()
    B, K, H, W = hm.shape
    idx = torch.empty((B, K), dtype=torch.int32, device=hm.device)
    conf = torch.empty((B, K), dtype=torch.float32, device=hm.device)
    check(lib.df3d_heatmap_argmax(_ptr(hm), 0 if hm.dtype == torch.float32 else 1, B, K, H, W,
                                  _ptr(idx), _ptr(conf), _stream()))
    return idx, conf


def heatmap_argmax_nhwc(hm, K):
    """(B,H,W,Cpad) float32 CUDA, decode channels [0,K)."""
    _need_cuda(hm)
    hm = hm.contiguous()
    B, H, W, Cp = hm.shape
    idx = torch.empty((B, K), dtype=torch.int32, device=hm.device)
    conf = torch.empty((B, K), dtype=torch.float32, device=hm.device)
    check(lib.df3d_heatmap_argmax_nhwc(_ptr(hm), B, H, W, Cp, K, _ptr(idx), _ptr(conf), _stream()))
    return idx, conf


def resize_gray_u8(images, size_hw):
    """images (B,Hs,Ws) uint8 on the device -> (B,Hd,Wd) uint8, bit-identical to
    cv2.resize(img, (Wd, Hd), interpolation=cv2.INTER_LINEAR) per image (the loader's host path)."""
    _need_cuda(images)
    if images.dtype != torch.uint8 or images.dim() != 3:
        raise ValueError("resize_gray_u8: expected a (B, H, W) uint8 tensor")
    images = images.contiguous()
    B, Hs, Ws = images.shape
    Hd, Wd = int(size_hw[0]), int(size_hw[1])
    out = torch.empty((B, Hd, Wd), dtype=torch.uint8, device=images.device)
    check(lib.df3d_resize_gray_u8(_ptr(images), B, Hs, Ws, _ptr(out), Hd, Wd, _stream()))
    return out


class JpegDecoder:
    """Device-side JPEG decode (luminance) through nvJPEG, bound at run time.  Opt-in: a few grey levels away
    from the host's libjpeg read, see include/df3d_b200.h.  backend: None = the GPU's hardware JPEG engines when
    the box has them, else nvJPEG's default; "hardware" / "default" ask for one (RuntimeError if absent)."""

    BACKENDS = {None: -1, "hardware": 3, "gpu_hybrid": 2, "default": 0}

    def __init__(self, backend=None):
        self._h = C.c_void_p()
        check(lib.df3d_jpeg_create_backend(C.byref(self._h), self.BACKENDS[backend]))

    @property
    def backend(self):
        return {3: "hardware", 2: "gpu_hybrid", 0: "default"}[lib.df3d_jpeg_backend(self._h)]

    def close(self):
        if self._h:
            lib.df3d_jpeg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def image_size(self, data):
        """-> (H, W) of a compressed stream (bytes)."""
        w, h = C.c_int(), C.c_int()
        check(lib.df3d_jpeg_info(self._h, C.cast(C.c_char_p(data), C.c_void_p), len(data), C.byref(w), C.byref(h)))
        return h.value, w.value

    def decode_gray(self, streams, device="cuda"):
        """streams: list of bytes, all of one size -> (n, H, W) uint8 tensor on the device (work enqueued on the
        current stream; keep `streams` alive until it has run)."""
        n = len(streams)
        if n == 0:
            return torch.empty((0, 0, 0), dtype=torch.uint8, device=device)
        streams = [bytes(s) for s in streams]                      # no copy for bytes objects
        H, W = self.image_size(streams[0])
        out = torch.empty((n, H, W), dtype=torch.uint8, device=device)
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(s), C.c_void_p).value for s in streams])   # the bytes' own buffers
        lens = (C.c_size_t * n)(*[len(s) for s in streams])
        with torch.cuda.device(out.device):
            check(lib.df3d_jpeg_decode_gray(self._h, ptrs, lens, n, _ptr(out), H, W, _stream()))
        return out


def pack_points2d(idx, C_, T, heatmap_shape, camera_ordering, image_shape):
    """idx (C*T,K) int32 (camera-major) -> points2d (C,T,2K,2) f64 normalised (row,col),
    pts_xy (C,T,2K,2) f64 pixel (x,y).  image_shape = [W, H] like Core.image_shape."""
    _need_cuda(idx)
    idx = idx.contiguous()
    K = idx.shape[1]
    if idx.shape[0] != C_ * T:
        raise ValueError("pack_points2d: idx must have C*T rows")
    Hh, Wh = heatmap_shape
    order = (C.c_int * C_)(*[int(c) for c in camera_ordering])
    p2d = torch.empty((C_, T, 2 * K, 2), dtype=torch.float64, device=idx.device)
    pxy = torch.empty_like(p2d)
    check(lib.df3d_pack_points2d(_ptr(idx), C_, T, K, Hh, Wh, order, int(image_shape[0]), int(image_shape[1]),
                                 _ptr(p2d), _ptr(pxy), _stream()))
    return p2d, pxy


def projection_matrices(cam_rt, intr4):
    """cam_rt (C,6), intr4 (C,4) -> P (C,3,4), R (C,3,3)."""
    _need_cuda(cam_rt, intr4)
    Cn = cam_rt.shape[0]
    P = torch.empty((Cn, 3, 4), dtype=torch.float64, device=cam_rt.device)
    R = torch.empty((Cn, 3, 3), dtype=torch.float64, device=cam_rt.device)
    check(lib.df3d_projection_matrices(_ptr(cam_rt.contiguous()), _ptr(intr4.contiguous()), Cn, _ptr(P), _ptr(R), _stream()))
    return P, R


def triangulate_dlt(P, pts_xy):
    """P (C,3,4) f64, pts_xy (C,T,J,2) f64 pixel (x,y) -> (T,J,3) f64."""
    _need_cuda(P, pts_xy)
    if P.dtype != torch.float64 or pts_xy.dtype != torch.float64:
        raise ValueError("triangulate_dlt: float64 tensors required")
    P = P.contiguous()
    pts_xy = pts_xy.contiguous()
    Cn, T, J, _ = pts_xy.shape
    out = torch.empty((T, J, 3), dtype=torch.float64, device=pts_xy.device)
    check(lib.df3d_triangulate_dlt(_ptr(P), _ptr(pts_xy), Cn, T, J, _ptr(out), _stream()))
    return out


def ba_workspace(Cn, T, J, device):
    n = lib.df3d_bundle_adjust_workspace_bytes(Cn, T, J)
    return torch.empty(n + 256, dtype=torch.uint8, device=device)


def _aligned_ptr(ws):
    p = ws.data_ptr()
    a = (p + 255) & ~255
    return C.c_void_p(a), ws.numel() - (a - p)


def _decode_report(rep_bytes):
    raw = rep_bytes.cpu().numpy().tobytes()
    r = _lib.BAReport.from_buffer_copy(raw)
    return {"cost0": r.cost0, "cost": r.cost, "reg": r.reg, "iters": r.iters,
            "accepted": r.accepted, "n_obs": r.n_obs, "status": r.status, "lsmr_itn": r.lsmr_itn, "lsmr_istop": r.lsmr_istop}


def bundle_adjust(cam_rt, intr4, pts_xy, pts3d, max_iters=20, ftol=1e-4, xtol=1e-8, gtol=1e-8, workspace=None, solver="lsmr"):
    """In-place bundle adjustment (SciPy's trust-region-reflective iteration, see csrc/bundle_adjust.cu).
    cam_rt (C,6) and pts3d (T,J,3) are updated.

    Returns a uint8 CUDA tensor holding the df3d_ba_report (decode with ``ba_report``; reading it
    synchronises).  Every reduction runs in a fixed order: the same inputs give the same bits, which is
    what lets a frame-sharded multi-GPU run solve the gathered problem replicated (pipeline.py)."""
    _need_cuda(cam_rt, intr4, pts_xy, pts3d)
    for t in (cam_rt, intr4, pts_xy, pts3d):
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("bundle_adjust: contiguous float64 tensors required")
    Cn, T, J, _ = pts_xy.shape
    ws = workspace if workspace is not None else ba_workspace(Cn, T, J, cam_rt.device)
    wp, wn = _aligned_ptr(ws)
    if solver not in ("lsmr", "exact"):
        raise ValueError("solver must be 'lsmr' (SciPy's truncated LSMR step) or 'exact' (Schur-complement solve)")
    opts = _lib.BAOpts(int(max_iters), float(ftol), float(xtol), float(gtol), 1 if solver == "lsmr" else 0)
    rep = torch.zeros(C.sizeof(_lib.BAReport), dtype=torch.uint8, device=cam_rt.device)
    check(lib.df3d_bundle_adjust(_ptr(cam_rt), _ptr(intr4), _ptr(pts_xy), Cn, T, J, C.byref(opts), _ptr(pts3d),
                                 _ptr(rep), wp, wn, _stream()))
    return rep


def ba_sharded_plan(Cn, T, J, world):
    """-> (n_blocks, partials_offset_bytes, [doubles per block of pass 0..3]) or None when the blocks of points do not
    split evenly over `world` ranks (the caller then solves replicated)."""
    nb, off, pd = C.c_int(), C.c_size_t(), (C.c_int * 4)()
    rc = lib.df3d_ba_sharded_plan(int(Cn), int(T), int(J), int(world), C.byref(nb), C.byref(off), pd)
    if rc == _lib.DF3D_EUNSUPPORTED:
        return None
    check(rc)
    return nb.value, off.value, list(pd)


def bundle_adjust_sharded(cam_rt, intr4, pts_xy, pts3d, group=None, max_iters=20, ftol=1e-4, xtol=1e-8, gtol=1e-8, workspace=None,
                          ranks=None):
    """Frame-sharded bundle adjustment (exact solver): every rank of `group` calls this with the SAME cameras and the SAME
    gathered 2-D points; the per-point work of each pass is split by blocks of points between the ranks, the per-block
    partial sums are all-gathered (one `all_gather_into_tensor` per pass) and summed in the single-GPU order, so cam_rt
    comes out bit-identical to ``bundle_adjust(..., solver="exact")`` on one GPU -- at 1 / world of its per-point work.
    pts3d: only the points of this rank's blocks are updated.  `ranks` = (rank, world) without a process group runs the
    passes of all `world` ranks one after the other on this GPU (tests)."""
    import torch.distributed as dist

    _need_cuda(cam_rt, intr4, pts_xy, pts3d)
    for t in (cam_rt, intr4, pts_xy, pts3d):
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("bundle_adjust_sharded: contiguous float64 tensors required")
    Cn, T, J, _ = pts_xy.shape
    if group is not None:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        local = [rank]
    else:
        world = int(ranks) if ranks else 1
        rank, local = 0, list(range(world))
    plan = ba_sharded_plan(Cn, T, J, world)
    if plan is None:
        raise ValueError("bundle_adjust_sharded: the point blocks do not split evenly over the ranks")
    n_blocks, off, pass_doubles = plan
    ws = workspace if workspace is not None else ba_workspace(Cn, T, J, cam_rt.device)
    wp, wn = _aligned_ptr(ws)
    base = wp.value - ws.data_ptr() + off                       # byte offset of the partials inside the workspace tensor
    opts = _lib.BAOpts(int(max_iters), float(ftol), float(xtol), float(gtol), 0)
    rep = torch.zeros(C.sizeof(_lib.BAReport), dtype=torch.uint8, device=cam_rt.device)
    st = _stream()
    check(lib.df3d_ba_sharded_begin(_ptr(cam_rt), Cn, T, J, C.byref(opts), wp, wn, st))
    per = n_blocks // world
    for it in range(1, int(max_iters) + 1):
        for p in range(4):
            for r in local:
                check(lib.df3d_ba_sharded_pass(p, r, world, _ptr(intr4), _ptr(pts_xy), _ptr(pts3d), Cn, T, J, wp, wn, st))
            if group is not None and world > 1:
                region = ws[base:base + n_blocks * pass_doubles[p] * 8].view(torch.float64)
                mine = region[rank * per * pass_doubles[p]:(rank + 1) * per * pass_doubles[p]].clone()
                dist.all_gather_into_tensor(region, mine, group=group)
            # (without a process group all blocks were computed here: one finish that owns every point)
            fr, fw = (rank, world) if group is not None else (0, 1)
            check(lib.df3d_ba_sharded_finish(p, it, fr, fw, _ptr(pts3d), Cn, T, J, wp, wn, st))
    check(lib.df3d_ba_sharded_end(_ptr(cam_rt), Cn, T, J, _ptr(rep), wp, wn, st))
    return rep


def bundle_adjust_launches(max_iters, solver="lsmr"):
    opts = _lib.BAOpts(int(max_iters), 1e-4, 1e-8, 1e-8, 1 if solver == "lsmr" else 0)
    return int(lib.df3d_bundle_adjust_launches(C.byref(opts)))


def ba_report(rep):
    return _decode_report(rep)


def reprojection_error(cam_rt, intr4, pts_xy, pts3d):
    """Mean L2 pixel error over the used observations, as a 0-d CUDA tensor."""
    _need_cuda(cam_rt, intr4, pts_xy, pts3d)
    Cn, T, J, _ = pts_xy.shape
    out = torch.empty(2, dtype=torch.float64, device=cam_rt.device)
    check(lib.df3d_reprojection_error(_ptr(cam_rt.contiguous()), _ptr(intr4.contiguous()), _ptr(pts_xy.contiguous()),
                                      _ptr(pts3d.contiguous()), Cn, T, J, _ptr(out), _stream()))
    return out[0] / out[1]


_TEMPLATE_MEDIANS = {}


def template_medians(device, template=None):
    """(2, 30) float64 on the device: per half the 12 median bone lengths and the 6 x 3 median alignment-joint
    coordinates of the procrustes template (constants of the template, df3d/procrustes.py:38-48, 102-116)."""
    from .procrustes import bone_lengths, read_template_pose3d
    from .skeleton import ALIGN_IDX

    key = (str(device), None if template is None else id(template))
    if key not in _TEMPLATE_MEDIANS or template is not None:
        tmpl = read_template_pose3d() if template is None else np.asarray(template, dtype=np.float64)
        half = tmpl.shape[1] // 2
        rows = []
        for h in range(2):
            t = tmpl[:, h * half:(h + 1) * half]
            rows.append(np.concatenate([np.median(bone_lengths(t), axis=0), np.median(t[:, ALIGN_IDX], axis=0).ravel()]))
        val = torch.as_tensor(np.stack(rows), device=device)
        if template is not None:
            return val
        _TEMPLATE_MEDIANS[key] = val
    return _TEMPLATE_MEDIANS[key]


def procrustes_workspace(T, device):
    return torch.empty(lib.df3d_procrustes_workspace_bytes(int(T)) + 256, dtype=torch.uint8, device=device)


def procrustes(pts3d, template=None, workspace=None):
    """(T,38,3) float64 CUDA -> registered (T,38,3): df3d.procrustes.procrustes_seperate on the device (global
    medians by radix select).  Enqueued on the current stream, no host synchronisation."""
    _need_cuda(pts3d)
    if pts3d.dtype != torch.float64 or pts3d.dim() != 3 or pts3d.shape[2] != 3:
        raise ValueError("procrustes: expected a (T, J, 3) float64 tensor")
    pts3d = pts3d.contiguous()
    T, J, _ = pts3d.shape
    out = torch.empty_like(pts3d)
    if T == 0:
        return out
    ws = workspace if workspace is not None else procrustes_workspace(T, pts3d.device)
    wp, wn = _aligned_ptr(ws)
    check(lib.df3d_procrustes(_ptr(pts3d), T, J, _ptr(template_medians(pts3d.device, template)), _ptr(out), wp, wn, _stream()))
    return out


def one_euro_filter(pts, freq=100.0, mincutoff=0.1, beta=2.0, dcutoff=1.0, t_first=1):
    """(T, J, D) float64 CUDA -> One-Euro filtered tracks (df3d.signal_util.filter_batch defaults; the 2-D variant
    filter_batch_2d is mincutoff=0.0001, beta=30, t_first=0)."""
    _need_cuda(pts)
    if pts.dtype != torch.float64:
        raise ValueError("one_euro_filter: float64 tensor required")
    pts = pts.contiguous()
    T = pts.shape[0]
    n = int(pts.numel() // T) if T else 0
    out = torch.empty_like(pts)
    if pts.numel() == 0:
        return out
    check(lib.df3d_one_euro_filter(_ptr(pts), T, n, float(freq), float(mincutoff), float(beta), float(dcutoff), int(t_first),
                                   _ptr(out), _stream()))
    return out


def smooth_pose2d(points2d, window_size=20, std_thr=5.0):
    """(T, J, 2) float64 CUDA pixel tracks -> df3d.signal_util.smooth_pose2d (pad = window edge replication)."""
    _need_cuda(points2d)
    if points2d.dtype != torch.float64:
        raise ValueError("smooth_pose2d: float64 tensor required")
    points2d = points2d.contiguous()
    T = points2d.shape[0]
    n = int(points2d.numel() // T) if T else 0
    out = torch.empty_like(points2d)
    if points2d.numel() == 0:
        return out
    check(lib.df3d_smooth_pose2d(_ptr(points2d), T, n, int(window_size), float(std_thr), _ptr(out), _stream()))
    return out


def intr_to_vec4(intr):
    """(C,3,3) camera matrices -> (C,4) fx, fy, cx, cy."""
    intr = np.asarray(intr, dtype=np.float64)
    return np.stack([intr[:, 0, 0], intr[:, 1, 1], intr[:, 0, 2], intr[:, 1, 2]], axis=1)
