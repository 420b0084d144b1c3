"""GPU parity of the pyba-half kernels (arg-max, packing, DLT, bundle adjustment) through the
C ABI, against the CPU oracle and the reference's golden pickles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import argmax as oargmax
from oracle import geometry as g
from oracle import pack as opack


@pytest.fixture(scope="module")
def ops(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from deepfly3d_b200 import ops as _ops

    return _ops


def _cuda(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


# ------------------------------------------------------------------ arg-max: bit-exact
@pytest.mark.parametrize("shape", [(3, 19, 64, 128), (2, 19, 64, 64), (1, 5, 7, 9), (4, 1, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_heatmap_argmax_exact(ops, shape, dtype):
    gen = torch.Generator().manual_seed(1)
    hm = torch.randn(shape, generator=gen).to(dtype)
    hm[0, 0].fill_(0.25)                     # all-tie plane -> index 0
    if shape[2] * shape[3] > 8:
        hm[-1, -1].view(-1)[5] = 9.0
        hm[-1, -1].view(-1)[7] = 9.0         # tie -> first occurrence
    idx, conf = ops.heatmap_argmax(hm.cuda())
    ref_idx, ref_conf = oargmax.heatmap_argmax(hm.float().numpy())
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(conf.cpu().numpy(), ref_conf)


def test_heatmap_argmax_nhwc_exact(ops):
    gen = torch.Generator().manual_seed(2)
    hm = torch.randn((5, 64, 64, 32), generator=gen)
    hm[1, :, :, 3] = 1.5
    idx, conf = ops.heatmap_argmax_nhwc(hm.cuda(), 19)
    ref_idx, ref_conf = oargmax.heatmap_argmax(hm.permute(0, 3, 1, 2)[:, :19].contiguous().numpy())
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(conf.cpu().numpy(), ref_conf)


def test_heatmap_argmax_empty_batch(ops):
    idx, conf = ops.heatmap_argmax(torch.zeros((0, 19, 8, 8), device="cuda"))
    assert idx.shape == (0, 19) and conf.shape == (0, 19)


# ------------------------------------------------------------------ packing: bit-exact
@pytest.mark.parametrize("order", [[0, 1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1, 0], [0, 6, 5, 4, 3, 2, 1]])
def test_pack_points2d_exact(ops, order):
    rng = np.random.default_rng(3)
    T, K, Hh, Wh = 9, 19, 64, 128
    idx = rng.integers(0, Hh * Wh, size=(7, T, K)).astype(np.int32)
    idx[0, 0, 0] = 0            # arg-max in the corner -> (0,0): dropped by the visibility rule
    idx[5, 1, 2] = 77           # row 0
    p2d, pxy = ops.pack_points2d(_cuda(idx.reshape(7 * T, K), torch.int32), 7, T, (Hh, Wh), order, [960, 480])
    ref = opack.pack_points2d(opack.indices_to_points2d(idx, (Hh, Wh)), order)
    assert np.array_equal(p2d.cpu().numpy(), ref)
    assert np.array_equal(pxy.cpu().numpy(), g.to_pixels_xy(ref, [960, 480]))


def test_pack_golden_roundtrip(ops, golden):
    """Feeding the indices behind the golden 2-D points reproduces the golden array bit for bit."""
    p38 = golden["result_2d"]["points2d"]
    order = golden["result_2d"]["camera_ordering"]
    T = p38.shape[1]
    idx = np.zeros((7, T, 19), dtype=np.int32)
    for slot, cam in enumerate(order):
        src = p38[cam, :, :19] if slot < 3 else p38[cam, :, 19:].copy()
        if slot > 3:
            src[..., 1] = 1 - src[..., 1]
        if slot != 3:
            idx[cam] = np.round(src[..., 0] * 64).astype(np.int32) * 128 + np.round(src[..., 1] * 128).astype(np.int32)
    p2d, _ = ops.pack_points2d(_cuda(idx.reshape(7 * T, 19), torch.int32), 7, T, (64, 128), order, [960, 480])
    assert np.array_equal(p2d.cpu().numpy(), p38)


# ------------------------------------------------------------------ DLT
def test_triangulate_matches_golden(ops, golden):
    r3 = golden["result_3d"]
    pts_xy = g.to_pixels_xy(golden["result_2d"]["points2d"], [960, 480])
    P = g.projection_matrices(r3["R"], r3["tvec"], r3["intr"])
    X = ops.triangulate_dlt(_cuda(P), _cuda(pts_xy)).cpu().numpy()
    np.testing.assert_allclose(X, r3["points3d_wo_procrustes"], atol=1e-9)   # reference test: 1e-5
    np.testing.assert_allclose(X, g.triangulate_dlt(P, pts_xy), atol=1e-9)


def test_triangulate_ragged_views(ops, golden):
    """0, 1, 2 ... 7 views per joint, zero coordinates in either axis, T not a multiple of the block."""
    rng = np.random.default_rng(4)
    c = golden["calib"]
    P = g.projection_matrices(c["R"], c["tvec"], c["intr"])
    T, J = 37, 38
    X = rng.normal(scale=1.5, size=(T * J, 3))
    pts = np.stack([g.project(X, c["R"][k], c["tvec"][k], c["intr"][k]) for k in range(7)]).reshape(7, T, J, 2)
    pts += rng.normal(scale=0.5, size=pts.shape)
    drop = rng.random((7, T, J)) < 0.45
    pts[drop] = 0.0
    pts[0, 0, 0] = (0.0, 13.0)              # one zero coordinate is enough to drop the view
    pts[1, 0, 1] = (17.0, 0.0)
    out = ops.triangulate_dlt(_cuda(P), _cuda(pts)).cpu().numpy()
    ref = g.triangulate_dlt(P, pts)
    few = g.visibility(pts).sum(0) < 2
    assert few.any() and np.all(out[few] == 0.0)
    np.testing.assert_allclose(out, ref, rtol=1e-8, atol=1e-8)


def test_projection_matrices(ops, golden):
    c = golden["calib"]
    cam = np.stack([np.concatenate([g.rodrigues_inv(c["R"][k]), c["tvec"][k]]) for k in range(7)])
    from deepfly3d_b200.ops import intr_to_vec4

    P, R = ops.projection_matrices(_cuda(cam), _cuda(intr_to_vec4(c["intr"])))
    np.testing.assert_allclose(R.cpu().numpy(), c["R"], atol=1e-13)
    np.testing.assert_allclose(P.cpu().numpy(), g.projection_matrices(c["R"], c["tvec"], c["intr"]), rtol=1e-13, atol=1e-9)


# ------------------------------------------------------------------ bundle adjustment
def _run_ba(ops, calib, pts_xy, **kw):
    from deepfly3d_b200.ops import intr_to_vec4

    cam0 = np.stack([np.concatenate([g.rodrigues_inv(calib["R"][k]), calib["tvec"][k]]) for k in range(7)])
    cam = _cuda(cam0)
    intr4 = _cuda(intr_to_vec4(calib["intr"]))
    pxy = _cuda(pts_xy)
    P0, _ = ops.projection_matrices(cam, intr4)
    X = ops.triangulate_dlt(P0, pxy)
    rep = ops.bundle_adjust(cam, intr4, pxy, X, **kw)
    P1, R1 = ops.projection_matrices(cam, intr4)
    X1 = ops.triangulate_dlt(P1, pxy)
    err = ops.reprojection_error(cam, intr4, pxy, X1)
    return cam.cpu().numpy(), R1.cpu().numpy(), X1.cpu().numpy(), ops.ba_report(rep), float(err), cam0


def test_bundle_adjust_golden(ops, golden):
    """Same input as the reference's test_calibration; north-star tolerance: 1e-3 mm on 3-D joints."""
    r3 = golden["result_3d"]
    pts_xy = g.to_pixels_xy(golden["result_2d"]["points2d"], [960, 480])
    cam, R1, X1, rep, err, cam0 = _run_ba(ops, golden["calib"], pts_xy)
    assert rep["n_obs"] == 1590
    assert abs(rep["cost0"] - 11953.29) < 0.05 and abs(rep["cost"] - 11136.13) < 0.05   # SciPy: 11953.29 -> 11136.13
    assert rep["status"] == 1 and rep["accepted"] <= 6
    assert np.abs(X1 - r3["points3d_wo_procrustes"]).max() < 1e-3
    assert np.abs(R1 - r3["R"]).max() < 1e-3
    assert abs(err - 2.942) < 5e-3
    # camera 3 has no observations: returned bit-identical
    assert np.array_equal(cam[3], cam0[3])


def test_bundle_adjust_vs_oracle_synthetic(ops, golden):
    """Perturbed cameras, synthetic skeleton, 40 frames: GPU LM vs the SciPy-TRF oracle."""
    rng = np.random.default_rng(5)
    c = golden["calib"]
    T, J = 40, 38
    tmpl = golden["template"]["points3d"]
    X = tmpl[rng.integers(0, tmpl.shape[0], size=T)] + rng.normal(scale=0.05, size=(T, J, 3))
    pts = np.zeros((7, T, J, 2))
    for k in (0, 1, 2):
        pts[k, :, :19] = g.project(X[:, :19].reshape(-1, 3), c["R"][k], c["tvec"][k], c["intr"][k]).reshape(T, 19, 2)
    for k in (4, 5, 6):
        pts[k, :, 19:] = g.project(X[:, 19:].reshape(-1, 3), c["R"][k], c["tvec"][k], c["intr"][k]).reshape(T, 19, 2)
    pts += rng.normal(scale=1.0, size=pts.shape) * (pts != 0)
    calib = {k: v.copy() for k, v in c.items()}
    for k in range(7):
        calib["R"][k] = g.rodrigues(g.rodrigues_inv(c["R"][k]) + rng.normal(scale=0.005, size=3))
        calib["tvec"][k] = c["tvec"][k] + rng.normal(scale=0.2, size=3)
    cam, R1, X1, rep, err, _ = _run_ba(ops, calib, pts)
    Ro, to = g.bundle_adjust(calib["R"], calib["tvec"], calib["intr"], pts)
    Xo = g.triangulate_dlt(g.projection_matrices(Ro, to, calib["intr"]), pts)
    erro = g.reprojection_error(Ro, to, calib["intr"], pts, Xo)
    assert rep["status"] == 1
    assert abs(err - erro) < 2e-3 * erro           # same optimum
    assert np.abs(X1 - Xo).max() < 1e-2            # gauge freedom: looser than on the golden case


def test_bundle_adjust_stepwise_equals_monolithic(ops, golden):
    from deepfly3d_b200.ops import intr_to_vec4

    pts_xy = g.to_pixels_xy(golden["result_2d"]["points2d"], [960, 480])
    c = golden["calib"]
    cam0 = np.stack([np.concatenate([g.rodrigues_inv(c["R"][k]), c["tvec"][k]]) for k in range(7)])
    intr4 = _cuda(intr_to_vec4(c["intr"]))
    pxy = _cuda(pts_xy)
    outs = []
    for fn in (ops.bundle_adjust, ops.bundle_adjust_distributed):
        cam = _cuda(cam0)
        P0, _ = ops.projection_matrices(cam, intr4)
        X = ops.triangulate_dlt(P0, pxy)
        fn(cam, intr4, pxy, X)
        outs.append((cam.cpu().numpy(), X.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_errors_are_reported(ops):
    from deepfly3d_b200._lib import Df3dError

    with pytest.raises(Df3dError):
        ops.triangulate_dlt(torch.zeros((9, 3, 4), dtype=torch.float64, device="cuda"),
                            torch.zeros((9, 1, 2, 2), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        ops.triangulate_dlt(torch.zeros((7, 3, 4), dtype=torch.float64), torch.zeros((7, 1, 2, 2), dtype=torch.float64))
