"""Procrustes registration of the triangulated skeleton to the template pose.

Host-side numpy, like the reference's own in-repo post-processing (df3d/procrustes.py:51-263,
df3d/plot_util.py:85-91): it needs global medians over all frames, so it runs once on the
gathered (T,38,3) array -- O(T) work next to the per-image network.  Left joints 0-18 and right
joints 19-37 are registered separately:
  scale  = median over 12 bones of  median_t(template bone length) / median_t(predicted length)
  centre = subtract the median of all points, multiply by scale
  rigid  = orthogonal Procrustes (reflection allowed, no scaling) of the median BODY_COXA /
           COXA_FEMUR joints onto the template's.
"""
import os

import numpy as np

from .skeleton import ALIGN_IDX, N_LEGS

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def read_template_pose3d(path=None):
    """Template skeleton (Tt,38,3): the `points3d` of the reference's data/df3d_result.pkl."""
    path = path or os.path.join(_DATA, "template.npz")
    with np.load(path) as z:
        return z["points3d"].astype(np.float64)


def bone_lengths(pts):
    legs = pts[:, : 5 * N_LEGS].reshape(pts.shape[0], N_LEGS, 5, 3)
    return np.linalg.norm(np.diff(legs, axis=2), axis=-1).reshape(pts.shape[0], N_LEGS * 4)


def orthogonal_fit(target, source):
    """Rotation/reflection Q and translation c minimising |source @ Q + c - target|."""
    mt, ms = target.mean(axis=0), source.mean(axis=0)
    t0, s0 = target - mt, source - ms
    t0 /= np.linalg.norm(t0)
    s0 /= np.linalg.norm(s0)
    U, _, Vt = np.linalg.svd(t0.T @ s0, full_matrices=False)
    Q = Vt.T @ U.T
    return Q, mt - ms @ Q


def register_half(pts, template):
    scale = np.median(np.median(bone_lengths(template), axis=0) / np.median(bone_lengths(pts), axis=0))
    centred = (pts - np.median(pts.reshape(-1, 3), axis=0)) * scale
    Q, c = orthogonal_fit(np.median(template[:, ALIGN_IDX], axis=0), np.median(centred[:, ALIGN_IDX], axis=0))
    return centred @ Q + c


def procrustes_seperate(pts, template=None):
    """(T,38,3) -> (T,38,3); name kept from the reference (df3d/procrustes.py:51)."""
    pts = np.asarray(pts, dtype=np.float64)
    template = read_template_pose3d() if template is None else template
    out = np.zeros_like(pts)
    half = pts.shape[1] // 2
    out[:, :half] = register_half(pts[:, :half], template[:, :half])
    out[:, half:] = register_half(pts[:, half:], template[:, half:])
    return out
