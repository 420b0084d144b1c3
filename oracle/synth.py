"""Synthetic multi-view geometry for the parity tests and the CPU baseline.

Test infrastructure (see ``oracle/__init__.py``).  Recipe: the *true* cameras are the packaged
initial calibration (reference ``data/calib.pkl``, loaded at ``df3d/core.py:234-239``) perturbed by
rvec ~ N(0, 0.01), tvec ~ N(0, 0.5); the ground-truth skeleton is the procrustes template
(``data/df3d_result.pkl['points3d']``, ``df3d/procrustes.py:38-48``) jittered by N(0, 0.05) per
frame; every joint is projected into the cameras that see it and quantised to the heat-map grid --
exactly what a hard arg-max read-out hands to ``Core.pose2d_estimation`` (``df3d/core.py:187-203``).
Bundle adjustment then starts from the *unperturbed* calibration like ``Core.calibrate_calc``
(``df3d/core.py:229-250``).
"""
import os

import numpy as np

from . import geometry as g
from . import pack as opack

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_calib():
    with np.load(os.path.join(GOLDEN, "calib.npz")) as z:
        return {k: z[k].astype(np.float64) for k in ("R", "tvec", "intr", "distort")}


def load_template():
    with np.load(os.path.join(GOLDEN, "template.npz")) as z:
        return z["points3d"].astype(np.float64)


def config3_points2d(T, seed=2, heatmap_shape=(64, 128), image_shape=(960, 480), rvec_sigma=0.01, tvec_sigma=0.5,
                     jitter=0.05, camera_ordering=range(7)):
    """-> (calib0, points2d (7,T,38,2) normalised (row, col) on the heat-map grid, X_true (T,38,3),
    true cameras {R, tvec}).  calib0 is the start point of the bundle adjustment."""
    rng = np.random.default_rng(seed)
    calib = load_calib()
    tmpl = load_template()
    Hh, Wh = heatmap_shape
    W, H = image_shape
    Rt = np.stack([g.rodrigues(g.rodrigues_inv(calib["R"][c]) + rng.normal(scale=rvec_sigma, size=3)) for c in range(7)])
    tt = calib["tvec"] + rng.normal(scale=tvec_sigma, size=(7, 3))
    X = tmpl[rng.integers(0, tmpl.shape[0], size=T)] + rng.normal(scale=jitter, size=(T, 38, 3))
    order = list(camera_ordering)
    p19 = np.zeros((7, T, 19, 2))
    for slot, c in enumerate(order):
        half = slice(0, 19) if slot < 4 else slice(19, 38)
        uv = g.project(X[:, half].reshape(-1, 3), Rt[c], tt[c], calib["intr"][c]).reshape(T, 19, 2)
        col = np.clip(np.round(uv[..., 0] / W * Wh), 1, Wh - 1)     # arg-max index on the heat-map grid
        row = np.clip(np.round(uv[..., 1] / H * Hh), 1, Hh - 1)
        if slot > 3:                                                # these cameras see the mirrored image
            col = Wh - col
        p19[c, ..., 0], p19[c, ..., 1] = row / Hh, col / Wh
    p38 = opack.pack_points2d(p19, order)
    return calib, p38, X, {"R": Rt, "tvec": tt}


def config3_geometry(T, seed=2, **kw):
    """-> (calib0, pts_xy (7,T,38,2) pixel (x, y), X_true)."""
    image_shape = kw.get("image_shape", (960, 480))
    calib, p38, X, _ = config3_points2d(T, seed=seed, **kw)
    return calib, g.to_pixels_xy(p38, list(image_shape)), X
