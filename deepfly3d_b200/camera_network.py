"""pyba-compatible CameraNetwork backed by the CUDA kernels.

Mirrors the part of ``pyba.CameraNetwork.CameraNetwork`` that ``df3d.core.Core`` touches
(reference call sites df3d/core.py:120-126, 165, 246-250, 295, 317, 355-360, 381): constructor
``CameraNetwork(points2d, calib, image_path, colors=, bones=)``, ``cam_list``, ``points2d``,
``points3d``, ``triangulate()``, ``bundle_adjust(update_intrinsic, update_distort)``,
``reprojection_error()``, ``summarize()``, ``has_calibration()``, ``__getitem__``.

numpy at the surface (like pyba), CUDA tensors inside.  The overlay/plot helpers of pyba
(``plot_2d``, ``get_image``) are visualisation and out of scope.
"""
import numpy as np
import torch

from . import ops
from .ops import intr_to_vec4


def rodrigues_vec(R):
    """Rotation matrix -> Rodrigues vector (angle in [0, pi]); host-side, 7 tiny matrices."""
    R = np.asarray(R, dtype=np.float64)
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(w)
    c = np.clip(0.5 * (np.trace(R) - 1.0), -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        # angle ~ pi: take the axis from the symmetric part
        d = np.sqrt(np.maximum((np.diag(R) + 1.0) * 0.5, 0.0))
        d[1] *= 1.0 if R[0, 1] >= 0 else -1.0
        d[2] *= 1.0 if R[0, 2] >= 0 else -1.0
        if abs(d[0]) < abs(d[1]) and abs(d[0]) < abs(d[2]) and (R[1, 2] > 0) != (d[1] * d[2] > 0):
            d[2] = -d[2]
        return d * (theta / np.linalg.norm(d))
    return w * (0.5 * theta / s)


class Camera:
    """One view: calibration + its (T,J,2) pixel observations (same axis order as passed in)."""

    def __init__(self, cam_id, points2d, R, tvec, intr, distort):
        self.cam_id = cam_id
        self.points2d = points2d
        self.R = np.array(R, dtype=np.float64)
        self.tvec = np.array(tvec, dtype=np.float64).reshape(3)
        self.intr = np.array(intr, dtype=np.float64)
        self.distort = np.array(distort, dtype=np.float64)

    def __getitem__(self, img_id):
        return self.points2d[img_id]

    @property
    def rvec(self):
        return rodrigues_vec(self.R)

    @property
    def P(self):
        return self.intr @ np.concatenate([self.R, self.tvec.reshape(3, 1)], axis=1)

    def is_empty(self):
        return not np.any(self.points2d)


class CameraNetwork:
    def __init__(self, points2d, calib=None, image_path=None, colors=None, bones=None, device="cuda"):
        """points2d: (C,T,J,2) in pixels, axis order (row*H, col*W) as built at core.py:247.
        calib: {cam_id: {R, tvec, intr, distort}} (extra keys such as 'meta' / result arrays ignored)."""
        if not torch.cuda.is_available():
            raise RuntimeError("CameraNetwork needs a CUDA device; there is no CPU fallback")
        self.points2d = np.ascontiguousarray(points2d, dtype=np.float64)
        self.image_path = image_path
        self.colors, self.bones = colors, bones
        self.device = torch.device(device)
        C = self.points2d.shape[0]
        self.cam_list = []
        for c in range(C):
            k = calib[c] if calib is not None and c in calib else None
            if k is None:
                self.cam_list.append(Camera(c, self.points2d[c], np.eye(3), np.zeros(3), np.eye(3), np.zeros(5)))
            else:
                self.cam_list.append(Camera(c, self.points2d[c], k["R"], k["tvec"], k["intr"], k["distort"]))
        self._calibrated = calib is not None
        # the kernels implement the reference's call -- pin-hole, update_distort=False on an all-zero distortion
        # (data/calib.pkl) and no skew; a calibration that carries either would be silently mis-projected
        for cam in self.cam_list:
            if np.any(cam.distort != 0) or cam.intr[0, 1] != 0:
                raise NotImplementedError(
                    f"camera {cam.cam_id}: non-zero lens distortion / skew is not supported by the CUDA projection "
                    "(the reference's packaged calibration has none, df3d/core.py:234-250)")
        self.points3d = None
        # (x, y) = (col, row) pixel coordinates on the device, the layout the kernels take
        self._pts_xy = torch.as_tensor(self.points2d[..., ::-1].copy(), device=self.device)
        if self._calibrated:
            self.triangulate()

    def __getitem__(self, cam_id):
        return self.cam_list[cam_id]

    def has_calibration(self):
        return self._calibrated

    # ---------------------------------------------------------------- device views
    def _cam_rt(self):
        rt = np.stack([np.concatenate([c.rvec, c.tvec]) for c in self.cam_list])
        return torch.as_tensor(rt, device=self.device)

    def _intr4(self):
        return torch.as_tensor(intr_to_vec4(np.stack([c.intr for c in self.cam_list])), device=self.device)

    def _P(self):
        return torch.as_tensor(np.stack([c.P for c in self.cam_list]), device=self.device)

    # ---------------------------------------------------------------- pyba surface
    def triangulate(self):
        """DLT over every (frame, joint) seen by >= 2 cameras (x != 0 and y != 0)."""
        X = ops.triangulate_dlt(self._P(), self._pts_xy)
        self.points3d = X.cpu().numpy()
        return self.points3d

    def bundle_adjust(self, update_intrinsic=False, update_distort=False, max_iters=20, ftol=1e-4, solver="lsmr"):
        """solver="lsmr" (default) reproduces the truncated LSMR step of the SciPy solver pyba calls and meets the
        reference test's own tolerances (tests/test_df3d.py:225-240); solver="exact" solves the same regularised
        Gauss-Newton step exactly (1-5e-5 mm away on the joints, ~4x faster)."""
        if update_intrinsic or update_distort:
            raise NotImplementedError(
                "only the reference's call bundle_adjust(update_intrinsic=False, update_distort=False) "
                "(df3d/core.py:249) is implemented")
        cam = self._cam_rt()
        cam0 = cam.clone()
        intr4 = self._intr4()
        P0, _ = ops.projection_matrices(cam, intr4)
        X = ops.triangulate_dlt(P0, self._pts_xy)
        rep = ops.bundle_adjust(cam, intr4, self._pts_xy, X, max_iters=max_iters, ftol=ftol, solver=solver)
        _, R = ops.projection_matrices(cam, intr4)
        cam_h, cam0_h, R_h = cam.cpu().numpy(), cam0.cpu().numpy(), R.cpu().numpy()
        for c, camera in enumerate(self.cam_list):
            if np.array_equal(cam_h[c], cam0_h[c]):
                continue                      # no observations: keep the stored matrices bit-identical
            camera.R = R_h[c]
            camera.tvec = cam_h[c, 3:].copy()
        self.points3d = X.cpu().numpy()
        self.ba_report = ops.ba_report(rep)
        return self.ba_report

    def reprojection_error(self):
        if self.points3d is None:
            self.triangulate()
        X = torch.as_tensor(self.points3d, device=self.device)
        return float(ops.reprojection_error(self._cam_rt(), self._intr4(), self._pts_xy, X))

    def summarize(self):
        return {c.cam_id: {"R": c.R, "tvec": c.tvec, "intr": c.intr, "distort": c.distort} for c in self.cam_list}
