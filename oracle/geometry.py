"""Oracle for the pyba half of the path: DLT triangulation + bundle adjustment.

Test infrastructure (see ``oracle/__init__.py``).  pyba itself is not vendored
in ``/root/reference``; this restates its behaviour as observed through the
reference's call sites and goldens:

* call sites: ``df3d/core.py:246-250`` (CameraNetwork + bundle_adjust),
  ``df3d/core.py:355-360`` (triangulate + summarize);
* golden pin: ``tests/data/reference_df3d/df3d_result_3d.pkl`` via
  ``tests/test_df3d.py:198-244`` (3-D atol 1e-5, cameras atol 1e-4).
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.sparse import lil_matrix


# ----------------------------------------------------------------------------
# Rodrigues (same convention as cv2.Rodrigues, used by pyba for the BA unknowns)
# ----------------------------------------------------------------------------
def rodrigues(rvec):
    """Rotation vector (3,) -> rotation matrix (3,3)."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = np.linalg.norm(rvec)
    if theta < 1e-300:
        return np.eye(3)
    k = rvec / theta
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return np.eye(3) * np.cos(theta) + (1.0 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * K


def rodrigues_inv(R):
    """Rotation matrix (3,3) -> rotation vector (3,) (angle in [0, pi])."""
    R = np.asarray(R, dtype=np.float64)
    # project on SO(3) the way OpenCV does (SVD clean-up) so tiny non-orthogonality
    # of a stored matrix does not leak into the angle
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt(np.dot(r, r) * 0.25)
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        ry = np.sqrt(max(t, 0.0)) * (1.0 if R[0, 1] >= 0 else -1.0)
        t = (R[2, 2] + 1) * 0.5
        rz = np.sqrt(max(t, 0.0)) * (1.0 if R[0, 2] >= 0 else -1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (theta / np.linalg.norm(v))
    return r * (0.5 * theta / s)


# ----------------------------------------------------------------------------
# pin-hole projection without distortion (cv2.projectPoints with zeros(5))
# ----------------------------------------------------------------------------
def project(X, R, tvec, intr):
    """X (N,3) world -> (N,2) pixel (x, y)."""
    Xc = X @ R.T + tvec
    x = Xc[:, 0] / Xc[:, 2]
    y = Xc[:, 1] / Xc[:, 2]
    return np.stack([intr[0, 0] * x + intr[0, 2], intr[1, 1] * y + intr[1, 2]], axis=1)


def projection_matrices(R, tvec, intr):
    """P_c = intr_c @ [R_c | t_c]  ->  (C,3,4)   (SURVEY Appendix B step 3)."""
    C = R.shape[0]
    P = np.zeros((C, 3, 4))
    for c in range(C):
        P[c] = intr[c] @ np.concatenate([R[c], tvec[c].reshape(3, 1)], axis=1)
    return P


def to_pixels_xy(points2d, image_shape):
    """Normalised (row, col) -> pixel (x, y).

    ``points2d * image_shape[::-1]`` (core.py:247; image_shape = [W, H]) gives
    (row*H, col*W); pyba then works in (x, y) = (col, row).
    """
    px = np.asarray(points2d, dtype=np.float64) * np.asarray(image_shape[::-1], dtype=np.float64)
    return px[..., ::-1].copy()


def visibility(pts_xy):
    """Observation is used iff neither pixel coordinate equals 0 (Appendix B step 2)."""
    return (pts_xy[..., 0] != 0) & (pts_xy[..., 1] != 0)


# ----------------------------------------------------------------------------
# DLT triangulation (CameraNetwork.triangulate, call site core.py:355)
# ----------------------------------------------------------------------------
def triangulate_dlt(P, pts_xy):
    """P (C,3,4), pts_xy (C,T,J,2) pixel (x,y) -> (T,J,3).

    For each (frame, joint) with >= 2 visible cameras stack, in ascending camera
    order, rows x*P[2]-P[0], y*P[2]-P[1]; X_h = last right singular vector.
    Joints with < 2 views stay 0.
    """
    C, T, J, _ = pts_xy.shape
    vis = visibility(pts_xy)
    out = np.zeros((T, J, 3))
    for t in range(T):
        for j in range(J):
            rows = []
            for c in range(C):
                if vis[c, t, j]:
                    x, y = pts_xy[c, t, j]
                    rows.append(x * P[c, 2] - P[c, 0])
                    rows.append(y * P[c, 2] - P[c, 1])
            if len(rows) >= 4:
                A = np.stack(rows)
                _, _, Vt = np.linalg.svd(A)
                Xh = Vt[-1]
                out[t, j] = Xh[:3] / Xh[3]
    return out


# ----------------------------------------------------------------------------
# bundle adjustment (CameraNetwork.bundle_adjust(update_intrinsic=False,
# update_distort=False), call site core.py:249)
# ----------------------------------------------------------------------------
def ba_observations(pts_xy):
    """Camera-major, then frame, then joint (Appendix B step 6)."""
    C, T, J, _ = pts_xy.shape
    vis = visibility(pts_xy)
    cam_idx, pt_idx, obs = [], [], []
    for c in range(C):
        tt, jj = np.nonzero(vis[c])
        cam_idx.append(np.full(tt.shape, c, dtype=np.int64))
        pt_idx.append(tt * J + jj)
        obs.append(pts_xy[c, tt, jj])
    return np.concatenate(cam_idx), np.concatenate(pt_idx), np.concatenate(obs)


def ba_residuals(x, n_cams, cam_idx, pt_idx, obs, intr):
    cam = x[: n_cams * 6].reshape(n_cams, 6)
    X = x[n_cams * 6 :].reshape(-1, 3)
    res = np.empty((obs.shape[0], 2))
    for c in range(n_cams):
        m = cam_idx == c
        if not m.any():
            continue
        res[m] = project(X[pt_idx[m]], rodrigues(cam[c, :3]), cam[c, 3:], intr[c]) - obs[m]
    return res.ravel()


def ba_sparsity(n_cams, n_points, cam_idx, pt_idx):
    m = cam_idx.size * 2
    n = n_cams * 6 + n_points * 3
    A = lil_matrix((m, n), dtype=int)
    i = np.arange(cam_idx.size)
    for s in range(6):
        A[2 * i, cam_idx * 6 + s] = 1
        A[2 * i + 1, cam_idx * 6 + s] = 1
    for s in range(3):
        A[2 * i, n_cams * 6 + pt_idx * 3 + s] = 1
        A[2 * i + 1, n_cams * 6 + pt_idx * 3 + s] = 1
    return A


def bundle_adjust(R, tvec, intr, pts_xy, return_info=False):
    """SciPy TRF + LSMR recipe (Appendix B steps 5-9).  Returns new (R, tvec).

    Unknowns: 7 x (rvec, tvec) ++ all 3-D points (initialised by DLT with the
    initial calibration).  Intrinsics / distortion are constants.
    """
    C, T, J, _ = pts_xy.shape
    X0 = triangulate_dlt(projection_matrices(R, tvec, intr), pts_xy)
    x0 = np.concatenate(
        [np.concatenate([rodrigues_inv(R[c]), tvec[c]]) for c in range(C)] + [X0.ravel()]
    )
    cam_idx, pt_idx, obs = ba_observations(pts_xy)
    A = ba_sparsity(C, T * J, cam_idx, pt_idx)
    sol = least_squares(
        ba_residuals,
        x0,
        jac_sparsity=A,
        x_scale="jac",
        ftol=1e-4,
        method="trf",
        args=(C, cam_idx, pt_idx, obs, intr),
    )
    cam = sol.x[: C * 6].reshape(C, 6)
    R_new = np.stack([rodrigues(cam[c, :3]) if (cam_idx == c).any() else R[c] for c in range(C)])
    t_new = np.stack([cam[c, 3:] if (cam_idx == c).any() else tvec[c] for c in range(C)])
    if return_info:
        return R_new, t_new, sol
    return R_new, t_new


def reprojection_error(R, tvec, intr, pts_xy, X):
    """Mean L2 pixel distance over all used observations (printed at core.py:250)."""
    C, T, J, _ = pts_xy.shape
    vis = visibility(pts_xy)
    errs = []
    Xf = X.reshape(-1, 3)
    for c in range(C):
        tt, jj = np.nonzero(vis[c])
        if tt.size == 0:
            continue
        p = project(Xf[tt * J + jj], R[c], tvec[c], intr[c])
        errs.append(np.linalg.norm(p - pts_xy[c, tt, jj], axis=1))
    return float(np.mean(np.concatenate(errs)))


def reorder_calib(calib, camera_ordering):
    """calib_reordered[cidx] = calib[idx] for (idx, cidx) in enumerate(ordering) (core.py:240-242)."""
    out = {k: np.empty_like(v) for k, v in calib.items()}
    for idx, cidx in enumerate(camera_ordering):
        for k in out:
            out[k][cidx] = calib[k][idx]
    return out


def calibrate_and_triangulate(points2d, calib, image_shape=(960, 480), camera_ordering=range(7)):
    """The whole 3-D half: core.py:229-250 then core.py:351-360 (without procrustes).

    points2d (7,T,38,2) normalised (row, col).  Returns dict with R, tvec (post-BA)
    and points3d_wo_procrustes (re-triangulated with the new cameras).
    """
    calib = reorder_calib(calib, list(camera_ordering))
    pts_xy = to_pixels_xy(points2d, list(image_shape))
    R, t = bundle_adjust(calib["R"], calib["tvec"], calib["intr"], pts_xy)
    X = triangulate_dlt(projection_matrices(R, t, calib["intr"]), pts_xy)
    return {"R": R, "tvec": t, "intr": calib["intr"], "distort": calib["distort"],
            "points3d_wo_procrustes": X, "pts_xy": pts_xy}
