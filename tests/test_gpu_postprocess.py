"""GPU parity of the post-processing kernels (csrc/postprocess.cu) through the C ABI: device procrustes against the
reference's golden points3d and the oracle, One-Euro filter / smooth_pose2d / the Core.get_points3d chain against
outputs of the reference's own code (tests/golden/signal.npz, see tests/golden/make_golden_signal.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import procrustes as oproc
from oracle import signal as osig

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ops(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from deepfly3d_b200 import ops as _ops

    return _ops


@pytest.fixture(scope="module")
def sig():
    with np.load(os.path.join(HERE, "golden", "signal.npz")) as z:
        return {k: z[k] for k in z.files}


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def test_procrustes_matches_golden(ops, golden, sig):
    """Reference test_calibration compares points3d at atol 1e-5 (tests/test_df3d.py:227-232); the device kernels
    reproduce the reference's own procrustes output to 1e-12 (exact medians, fp64 3x3 Jacobi SVD)."""
    r3 = golden["result_3d"]
    out = ops.procrustes(_cuda(r3["points3d_wo_procrustes"])).cpu().numpy()      # T = 15: odd counts
    np.testing.assert_allclose(out, r3["points3d"], rtol=0, atol=1e-12)
    out = ops.procrustes(_cuda(sig["pts3d_in"])).cpu().numpy()                   # T = 48: even counts
    np.testing.assert_allclose(out, sig["procrustes_out"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("T", [1, 2, 7, 1000, 4097])
def test_procrustes_vs_oracle_sizes(ops, golden, T):
    rng = np.random.default_rng(T)
    tmpl = golden["template"]["points3d"]
    base = golden["result_3d"]["points3d_wo_procrustes"]
    X = base[rng.integers(0, base.shape[0], size=T)] + rng.normal(scale=0.05, size=(T, 38, 3))
    X[rng.integers(0, T), 3] = X[0, 3]                       # duplicates among the values being selected
    out = ops.procrustes(_cuda(X)).cpu().numpy()
    np.testing.assert_allclose(out, oproc.procrustes_separate(X, tmpl), rtol=0, atol=1e-11)


def test_procrustes_other_template(ops, golden):
    rng = np.random.default_rng(3)
    tmpl = golden["template"]["points3d"] * 1.7 + 0.3
    X = golden["result_3d"]["points3d_wo_procrustes"] + rng.normal(scale=0.02, size=(15, 38, 3))
    out = ops.procrustes(_cuda(X), template=tmpl).cpu().numpy()
    np.testing.assert_allclose(out, oproc.procrustes_separate(X, tmpl), rtol=0, atol=1e-11)


def test_one_euro_bit_exact(ops, sig):
    """Same IEEE operations in the same order as df3d/signal_util.py:31-66 -> identical bits."""
    out = ops.one_euro_filter(_cuda(sig["pts3d_in"])).cpu().numpy()
    assert np.array_equal(out, sig["filter_batch_out"]), np.abs(out - sig["filter_batch_out"]).max()
    out = ops.one_euro_filter(_cuda(sig["pts3d_in"]), freq=30).cpu().numpy()
    assert np.array_equal(out, sig["filter_batch_freq30_out"])
    out = ops.one_euro_filter(_cuda(sig["pts2d_in"]), mincutoff=0.0001, beta=30, t_first=0).cpu().numpy()
    assert np.array_equal(out, sig["filter_batch_2d_out"]), np.abs(out - sig["filter_batch_2d_out"]).max()
    assert ops.one_euro_filter(torch.zeros((0, 38, 3), dtype=torch.float64, device="cuda")).shape == (0, 38, 3)
    with pytest.raises(RuntimeError):
        ops.one_euro_filter(_cuda(sig["pts3d_in"]), freq=0.0)              # "freq should be >0" (signal_util.py:33-34)


def test_smooth_pose2d_matches_reference(ops, sig):
    out = ops.smooth_pose2d(_cuda(sig["pts2d_in"])).cpu().numpy()
    np.testing.assert_allclose(out, sig["smooth_pose2d_out"], rtol=0, atol=1e-9)
    short = sig["pts2d_in"][:5]                                             # shorter than the window: edge replication
    np.testing.assert_allclose(ops.smooth_pose2d(_cuda(short)).cpu().numpy(), osig.smooth_pose2d(short), rtol=0, atol=1e-9)


def test_core_get_points3d_chain(ops, sig):
    """Core.get_points3d (df3d/core.py:332-343): procrustes -> normalize_pose_3d(rotate=True) -> filter_batch."""
    from deepfly3d_b200.core import Core

    class _Net:
        points3d = sig["pts3d_in"]

    core = object.__new__(Core)
    core.camNet = _Net()
    np.testing.assert_allclose(core.get_points3d(), sig["get_points3d_out"], rtol=0, atol=1e-11)
