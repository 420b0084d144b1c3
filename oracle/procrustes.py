"""Oracle restatement of the reference's procrustes registration.

Test infrastructure (see ``oracle/__init__.py``).  Follows
``df3d/procrustes.py:51-89`` (procrustes_seperate), ``:92-151`` (procrustes),
``:154-263`` (__procrustes, scaling=False, reflection='best') and
``df3d/plot_util.py:85-91`` (normalize_pose_3d: subtract the median of all
points).  Pinned by golden ``points3d`` (tests/test_df3d.py:227-232).
"""
import numpy as np

# joints of one 19-joint half whose type is BODY_COXA or COXA_FEMUR
# (df3d/skeleton_fly.py:16-55: three 5-joint legs, joint types repeat 0..4)
BODY_COXA_IDX = [0, 1, 5, 6, 10, 11]
N_LIMBS = 3


def _bone_lengths(p):
    """p (T,19,3) -> (T,12): 4 adjacent-joint distances for each of 3 legs."""
    legs = p[:, : 5 * N_LIMBS].reshape(p.shape[0], N_LIMBS, 5, 3)
    return np.linalg.norm(legs[:, :, 1:] - legs[:, :, :-1], axis=-1).reshape(p.shape[0], -1)


def _rigid_fit(X, Y):
    """Orthogonal map T and offset c with Y @ T + c ~ X (no scaling, reflection allowed)."""
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    X0 = X0 / np.sqrt((X0 ** 2).sum())
    Y0 = Y0 / np.sqrt((Y0 ** 2).sum())
    U, _, Vt = np.linalg.svd(X0.T @ Y0, full_matrices=False)
    T = Vt.T @ U.T
    return T, muX - muY @ T


def procrustes_half(pts, template):
    """pts (T,19,3), template (Tt,19,3) -> aligned (T,19,3)."""
    pts = pts.copy()
    s = np.median(np.median(_bone_lengths(template), axis=0) / np.median(_bone_lengths(pts), axis=0))
    pts -= np.median(pts.reshape(-1, 3), axis=0)
    pts *= s
    t_bc = np.median(template[:, BODY_COXA_IDX], axis=0)
    p_bc = np.median(pts[:, BODY_COXA_IDX], axis=0)
    T, c = _rigid_fit(t_bc, p_bc)
    return pts @ T + c


def procrustes_separate(pts, template):
    """pts (T,38,3), template (Tt,38,3): left joints 0-18 and right 19-37 separately."""
    out = np.zeros_like(pts)
    out[:, :19] = procrustes_half(pts[:, :19], template[:, :19])
    out[:, 19:38] = procrustes_half(pts[:, 19:38], template[:, 19:38])
    return out
