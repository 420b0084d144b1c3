"""Full-size (BASELINE.json configs[1] / configs[2]) checks through size-independent properties -- the
oracle cannot run 1 792 images of the 8-stack network in test time, so at that size the CUDA path is checked
against itself and against the arithmetic it must satisfy:
  * the 1 792-image forward equals the same images pushed through in small batches (images are independent:
    tile scheduling, CTA pairs, phantom tiles and chunking must not leak between images), and is reproducible;
  * the returned arg-max is the first-occurrence arg-max of the returned heat-map (oracle rule, README.md:404);
  * mirroring an image and setting its flip flag gives the un-mirrored image's result (reference df3d/core.py:179);
  * 1 000 frames (configs[2]): DLT of exact projections returns the points, bundle adjustment started at the
    optimum stays there, and its result does not depend on how many frames surround a frame block.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import argmax as oargmax
from oracle import geometry as g
from oracle import hourglass as ohg


@pytest.fixture(scope="module")
def mods(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from deepfly3d_b200 import hourglass, inference, ops

    return hourglass, inference, ops


def test_benchmark_batch_equals_small_batches(mods):
    hourglass, inference, _ = mods
    n = 7 * 256
    sd = inference.random_state_dict(8, seed=0)
    base = ohg.to_uint8(ohg.synthetic_images(64, 256, 256, seed=11))
    img = base.repeat(n // 64, 1, 1).contiguous()
    img[64:] = torch.roll(img[64:], shifts=3, dims=2)           # not all copies identical
    img = img.cuda()
    eng = hourglass.HourglassEngine(sd, 256, 256, max_batch=n)
    idx, conf, heat = eng.forward(img, return_heatmap=True)
    idx2, conf2 = eng.forward(img)
    torch.cuda.synchronize()
    assert torch.equal(idx, idx2) and torch.equal(conf, conf2), "forward is not reproducible"
    # arg-max of the returned heat-map (first occurrence), on a slice the CPU handles in seconds
    sl = slice(0, n, 97)
    hm = heat[sl, :, :, :eng.num_classes].permute(0, 3, 1, 2).contiguous().cpu().numpy()
    ref_idx, ref_conf = oargmax.heatmap_argmax(hm)
    assert np.array_equal(idx[sl].cpu().numpy(), ref_idx)
    assert np.array_equal(conf[sl].cpu().numpy(), ref_conf)
    # the same images in batches of 5 (odd: partially filled and phantom tiles on the low-resolution levels)
    small = hourglass.HourglassEngine(sd, 256, 256, max_batch=5)
    for start in (0, 891, n - 5):
        i5, c5 = small.forward(img[start:start + 5])
        torch.cuda.synchronize()
        assert torch.equal(i5, idx[start:start + 5]) and torch.equal(c5, conf[start:start + 5])
    eng.close()
    small.close()


def test_mirror_flag_equals_mirrored_input(mods):
    hourglass, inference, _ = mods
    sd = inference.random_state_dict(2, seed=1)
    img = ohg.to_uint8(ohg.synthetic_images(6, 256, 512, seed=12)).cuda()
    eng = hourglass.HourglassEngine(sd, 256, 512, max_batch=6)
    ia, ca = eng.forward(torch.flip(img, dims=[2]).contiguous(), flip=torch.ones(6, dtype=torch.uint8))
    ib, cb = eng.forward(img)
    torch.cuda.synchronize()
    assert torch.equal(ia, ib) and torch.equal(ca, cb)
    eng.close()


def _synthetic_views(golden, T, rng):
    c = golden["calib"]
    tmpl = golden["template"]["points3d"]
    X = tmpl[rng.integers(0, tmpl.shape[0], size=T)] + rng.normal(scale=0.05, size=(T, 38, 3))
    pts = np.zeros((7, T, 38, 2))
    for k in (0, 1, 2):
        pts[k, :, :19] = g.project(X[:, :19].reshape(-1, 3), c["R"][k], c["tvec"][k], c["intr"][k]).reshape(T, 19, 2)
    for k in (4, 5, 6):
        pts[k, :, 19:] = g.project(X[:, 19:].reshape(-1, 3), c["R"][k], c["tvec"][k], c["intr"][k]).reshape(T, 19, 2)
    return c, X, pts


def test_thousand_frames_geometry_properties(mods, golden):
    _, _, ops = mods
    from deepfly3d_b200.ops import intr_to_vec4

    rng = np.random.default_rng(21)
    T = 1000
    c, X, pts = _synthetic_views(golden, T, rng)
    cam0 = np.stack([np.concatenate([g.rodrigues_inv(c["R"][k]), c["tvec"][k]]) for k in range(7)])
    cam = torch.as_tensor(cam0).cuda()
    intr4 = torch.as_tensor(intr_to_vec4(c["intr"])).cuda()
    pxy = torch.as_tensor(pts).cuda()
    P, _ = ops.projection_matrices(cam, intr4)
    # exact projections triangulate back to the points (focal length 16 000 px: fp64 DLT, no normalisation)
    Xt = ops.triangulate_dlt(P, pxy)
    assert np.abs(Xt.cpu().numpy() - X).max() < 1e-6
    # bundle adjustment started at the optimum: nothing to do
    rep = ops.ba_report(ops.bundle_adjust(cam, intr4, pxy, Xt.clone()))
    assert rep["n_obs"] == 6 * 19 * T and rep["cost"] <= rep["cost0"] + 1e-12 and rep["cost"] < 1e-6
    moved = np.abs(cam.cpu().numpy() - cam0)
    assert moved.max() < 1e-6, moved.max()
    assert np.array_equal(cam.cpu().numpy()[3], cam0[3])      # the camera without observations is returned untouched
    # DLT is per (frame, joint): a block of frames gives the same points inside or outside the long sequence
    blk = slice(400, 464)
    Xb = ops.triangulate_dlt(P, pxy[:, blk].contiguous())
    assert torch.equal(Xb, Xt[blk])
