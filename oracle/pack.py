"""Oracle restatement of the 19 -> 38 joint packing of ``Core.pose2d_estimation``.

Test infrastructure (see ``oracle/__init__.py``).  Follows ``df3d/core.py:187-203`` statement by
statement, including the quirks the goldens confirm: camera ordering[3]
is dropped, unseen joints of the un-flipped cameras become (0, 1).
"""
import numpy as np


def indices_to_points2d(idx, heatmap_shape):
    """flat arg-max index (..., K) -> normalised (row/Hh, col/Wh) float64 (..., K, 2) (README.md:404)."""
    Hh, Wh = heatmap_shape
    idx = np.asarray(idx)
    return np.stack([(idx // Wh) / Hh, (idx % Wh) / Wh], axis=-1).astype(np.float64)


def pack_points2d(points2d, camera_ordering):
    """points2d (7,T,19,2) as returned by df2d -> (7,T,38,2) as stored by Core."""
    order = np.asarray(camera_ordering)
    p = np.asarray(points2d, dtype=np.float64)
    out = np.zeros((p.shape[0], p.shape[1], p.shape[2] * 2, 2))
    out[order[:3], :, :19] = p[order[:3]]          # core.py:190
    out[order[4:], :, 19:] = p[order[4:]]          # core.py:191
    out[order[2], :, 15:] = 0                      # core.py:194
    out[order[4], :, 19 + 15:] = 0                 # core.py:195
    for cidx in (4, 5, 6):                         # core.py:198-199
        out[order[cidx], ..., 1] = 1 - out[order[cidx], ..., 1]
    return out
