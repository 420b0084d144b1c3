"""Pins the oracle's restatement of the post-processing filters to outputs of the reference's own code
(tests/golden/signal.npz, written by tests/golden/make_golden_signal.py from df3d/signal_util.py,
df3d/plot_util.py, df3d/procrustes.py run in the build container)."""
import os

import numpy as np
import pytest

from oracle import procrustes as oproc
from oracle import signal as osig

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sig():
    with np.load(os.path.join(HERE, "golden", "signal.npz")) as z:
        return {k: z[k] for k in z.files}


def test_one_euro_matches_reference(sig):
    np.testing.assert_allclose(osig.one_euro_batch(sig["pts3d_in"]), sig["filter_batch_out"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(osig.one_euro_batch(sig["pts3d_in"], freq=30), sig["filter_batch_freq30_out"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(osig.one_euro_batch(sig["pts2d_in"], mincutoff=0.0001, beta=30, t_first=0),
                               sig["filter_batch_2d_out"], rtol=0, atol=1e-9)


def test_smooth_pose2d_matches_reference(sig):
    out = osig.smooth_pose2d(sig["pts2d_in"])
    np.testing.assert_allclose(out, sig["smooth_pose2d_out"], rtol=0, atol=1e-9)
    changed = np.abs(sig["smooth_pose2d_out"] - sig["pts2d_in"]) > 1e-9
    assert 0.05 < changed.mean() < 0.999        # both branches of the std threshold are exercised


def test_get_points3d_chain_matches_reference(sig, golden):
    np.testing.assert_allclose(osig.normalize_pose_3d(sig["pts3d_in"]), sig["normalize_rotate_out"], rtol=0, atol=1e-12)
    p = oproc.procrustes_separate(sig["pts3d_in"], golden["template"]["points3d"])
    np.testing.assert_allclose(p, sig["procrustes_out"], rtol=0, atol=1e-12)
    out = osig.one_euro_batch(osig.normalize_pose_3d(p))
    np.testing.assert_allclose(out, sig["get_points3d_out"], rtol=0, atol=1e-11)
