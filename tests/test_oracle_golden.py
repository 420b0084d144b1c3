"""Pins the 3-D half of the oracle to the reference's own golden pickles.

Same inputs and tolerances as the reference's test_calibration (tests/test_df3d.py:198-244):
points2d from df3d_result_2d.pkl + data/calib.pkl -> BA -> DLT -> procrustes, compared with
df3d_result_3d.pkl at atol 1e-5 (3-D) and 1e-4 (cameras).
"""
import numpy as np
import pytest

from oracle import geometry as g
from oracle import pack, procrustes


@pytest.fixture(scope="module")
def solved(golden):
    return g.calibrate_and_triangulate(golden["result_2d"]["points2d"], golden["calib"],
                                       image_shape=(960, 480), camera_ordering=golden["result_2d"]["camera_ordering"])


def test_known_answers(golden, solved):
    pts_xy = solved["pts_xy"]
    vis = g.visibility(pts_xy)
    assert int(vis.sum()) == 1590
    views = vis.sum(0).ravel()
    assert {int(v): int((views == v).sum()) for v in np.unique(views)} == {0: 0, 2: 120, 3: 450} or \
        {int(v): int((views == v).sum()) for v in np.unique(views)} == {2: 120, 3: 450}
    err = g.reprojection_error(solved["R"], solved["tvec"], solved["intr"], pts_xy, solved["points3d_wo_procrustes"])
    assert abs(err - 2.942) < 5e-3                      # "Reprojection error is ..." (core.py:250)


def test_dlt_with_golden_cameras(golden, solved):
    r3 = golden["result_3d"]
    P = g.projection_matrices(r3["R"], r3["tvec"], r3["intr"])
    X = g.triangulate_dlt(P, solved["pts_xy"])
    np.testing.assert_allclose(X, r3["points3d_wo_procrustes"], atol=1e-10)


def test_bundle_adjust_matches_golden(golden, solved):
    r3 = golden["result_3d"]
    np.testing.assert_allclose(solved["points3d_wo_procrustes"], r3["points3d_wo_procrustes"], atol=1e-5)
    np.testing.assert_allclose(solved["R"], r3["R"], atol=1e-4)
    np.testing.assert_allclose(solved["tvec"], r3["tvec"], atol=1e-4)
    np.testing.assert_allclose(solved["intr"], r3["intr"], atol=1e-4)
    np.testing.assert_allclose(solved["distort"], r3["distort"], atol=1e-4)
    # camera 3 has no observations and must come back bit-identical to calib.pkl
    assert np.array_equal(solved["R"][3], golden["calib"]["R"][3])
    assert np.array_equal(solved["tvec"][3], golden["calib"]["tvec"][3])


def test_procrustes_matches_golden(golden, solved):
    r3 = golden["result_3d"]
    tmpl = golden["template"]["points3d"]
    np.testing.assert_allclose(procrustes.procrustes_separate(r3["points3d_wo_procrustes"], tmpl), r3["points3d"], atol=1e-12)
    np.testing.assert_allclose(procrustes.procrustes_separate(solved["points3d_wo_procrustes"], tmpl), r3["points3d"], atol=1e-5)


def test_pack_reproduces_golden_layout(golden):
    """The golden (7,T,38,2) array must be a fixed point of unpack -> pack (core.py:187-203)."""
    p38 = golden["result_2d"]["points2d"]
    order = golden["result_2d"]["camera_ordering"]
    # invert the packing: per-camera 19-joint predictions in the (possibly mirrored) image frame
    p19 = np.zeros((7, p38.shape[1], 19, 2))
    for slot, cam in enumerate(order):
        if slot < 3:
            p19[cam] = p38[cam, :, :19]
        elif slot > 3:
            p19[cam] = p38[cam, :, 19:]
            p19[cam, ..., 1] = 1 - p19[cam, ..., 1]
    repacked = pack.pack_points2d(p19, order)
    np.testing.assert_array_equal(repacked, p38)
    # grid facts: 64 x 128 heat-map, hard arg-max
    nz = p38[p38[..., 0] != 0]
    assert np.all(nz[:, 0] * 64 == np.round(nz[:, 0] * 64)) and np.all(nz[:, 1] * 128 == np.round(nz[:, 1] * 128))


def test_rodrigues_roundtrip_against_opencv(golden):
    cv2 = pytest.importorskip("cv2")
    for c in range(7):
        R = golden["calib"]["R"][c]
        rv = g.rodrigues_inv(R)
        np.testing.assert_allclose(rv, cv2.Rodrigues(R)[0].ravel(), atol=1e-12)
        np.testing.assert_allclose(g.rodrigues(rv), cv2.Rodrigues(rv)[0], atol=1e-12)
    X = np.random.default_rng(0).normal(size=(20, 3))
    c = golden["calib"]
    ours = g.project(X, c["R"][0], c["tvec"][0], c["intr"][0])
    ref = cv2.projectPoints(X, cv2.Rodrigues(c["R"][0])[0], c["tvec"][0], c["intr"][0], np.zeros(5))[0].reshape(-1, 2)
    np.testing.assert_allclose(ours, ref, rtol=1e-12, atol=1e-8)
