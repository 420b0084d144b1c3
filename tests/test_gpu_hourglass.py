"""End-to-end parity of the CUDA hourglass (tcgen05 convs + fused epilogues + arg-max) against the
bf16-emulating CPU oracle on seeded weights and inputs.

Parity for this half is UNPINNED w.r.t. the reference (no pretrained weights, see
oracle/hourglass.py); what is pinned here is kernel == oracle arithmetic, against TWO oracles:
  * the bf16-emulating oracle rounds where the kernels round, so the only difference is the fp32 summation
    order inside a convolution, which can flip a bf16 rounding now and then: heat-map within 1 % of its
    dynamic range, arg-max identical wherever the oracle's peak is separated from the runner-up by more than
    that noise, confidence within the same bound;
  * the plain fp32 oracle -- no bf16, and the inter-stack re-injection x + fc_(y) + score_(score(y)) evaluated
    as the three separate convolutions of the published network, NOT the merged form the kernels (and the
    emulating oracle) use: a wrong merge formula or a misplaced rounding point would show here.  Bound: 2.5 % of
    the range (bf16 activations through up to 378 convolutions; the emulating oracle itself sits 0.6-0.75 %
    from fp32); the arg-max mismatch rate against fp32 is printed (hard part 2).
With seeded random weights the maps are noise-like, so many peaks are near ties; test_trained_network_*
below repeats the comparison on a network that was trained to produce peaked maps.  The reference's own test
demands +-1 heat-map row (tests/test_df3d.py:171: atol=0.02).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import argmax as oargmax
from oracle import hourglass as ohg


@pytest.fixture(scope="module")
def hgmod(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from deepfly3d_b200 import hourglass

    return hourglass


def _compare(hgmod, stacks, H, W, B, seed, flip=None, float_input=False):
    model = ohg.make_model(stacks, seed=seed)
    img = ohg.to_uint8(ohg.synthetic_images(B, H, W, seed=seed + 1))
    eng = hgmod.HourglassEngine(model.state_dict(), H, W, max_batch=max(B, 1))
    fl = None if flip is None else torch.tensor(flip, dtype=torch.uint8)
    x = ohg.preprocess_u8(img, flip=flip)
    if float_input:
        idx, conf, heat = eng.forward(ohg.preprocess_u8(img).cuda(), flip=None if fl is None else fl.cuda(), return_heatmap=True)
    else:
        idx, conf, heat = eng.forward(img.cuda(), flip=None if fl is None else fl.cuda(), return_heatmap=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = model(x, emulate_bf16=True, gray_fold=not float_input)[-1]   # (B,K,Hh,Wh) fp32
        ref32 = model(x)[-1]                                               # fp32 everywhere, unmerged re-injection
    got = heat[..., :19].permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    rng_ = (ref.max() - ref.min()).item()
    err = (got - ref).abs().max().item() / rng_
    err32 = (got - ref32).abs().max().item() / (ref32.max() - ref32.min()).item()
    idx32, _ = oargmax.heatmap_argmax(ref32.numpy())
    mis32 = float((idx.cpu().numpy() != idx32).mean())
    d32 = np.abs(np.stack(np.divmod(idx.cpu().numpy(), W // 4)) - np.stack(np.divmod(idx32, W // 4))).max(axis=0)
    print(f"  vs fp32 unmerged oracle: heat err {err32:.4f} of range, arg-max mismatch rate {mis32:.3f} "
          f"({float((d32 > 1).mean()):.3f} by more than one pixel)")
    assert err32 < 0.025, f"heat-map deviates from the fp32 (unmerged) oracle: {err32}"
    ref_idx, ref_conf = oargmax.heatmap_argmax(ref.numpy())
    # decode of the kernel's own heat-map must be bit-exact (integer index, fp32 peak)
    own_idx, own_conf = oargmax.heatmap_argmax(got.numpy())
    assert np.array_equal(idx.cpu().numpy(), own_idx)
    assert np.array_equal(conf.cpu().numpy(), own_conf)
    agree = float((idx.cpu().numpy() == ref_idx).mean())
    # where they disagree the two candidates must be a near tie in the oracle's map
    flat = ref.flatten(2).numpy()
    gi = idx.cpu().numpy()
    gap = np.take_along_axis(flat, ref_idx[..., None].astype(np.int64), -1)[..., 0] - \
        np.take_along_axis(flat, gi[..., None].astype(np.int64), -1)[..., 0]
    # integer-exact wherever the oracle's peak beats its runner-up by more than twice the
    # measured heat-map deviation (a flip would need both values to move against each other)
    top2 = np.sort(flat, axis=-1)[..., -2]
    margin = np.take_along_axis(flat, ref_idx[..., None].astype(np.int64), -1)[..., 0] - top2
    clear = margin > 2.0 * err * rng_
    assert np.array_equal(gi[clear], ref_idx[clear]), "arg-max differs on a well-separated peak"
    eng.close()
    print(f"  well-separated peaks: {clear.mean():.3f} of joints, all identical")
    return err, agree, float(gap.max() / rng_), float(np.abs(conf.cpu().numpy() - ref_conf).max() / rng_)


def test_two_stack_reference_shape(hgmod):
    """config 1 shape: 2 stacks, 256 x 512 input -> 64 x 128 heat-map, mirrored cameras included."""
    err, agree, gap, cerr = _compare(hgmod, 2, 256, 512, 3, seed=0, flip=[False, True, True])
    print(f"2-stack 256x512: heat err {err:.4f} of range, arg-max agreement {agree:.3f}, worst gap {gap:.4f}")
    assert err < 0.01 and gap < 0.01 and cerr < 0.01
    assert agree > 0.7


def test_eight_stack_benchmark_shape(hgmod):
    """config 2 shape: 8 stacks, 256 x 256 input -> 64 x 64 heat-map."""
    err, agree, gap, cerr = _compare(hgmod, 8, 256, 256, 2, seed=1)
    print(f"8-stack 256x256: heat err {err:.4f} of range, arg-max agreement {agree:.3f}, worst gap {gap:.4f}")
    assert err < 0.01 and gap < 0.01 and cerr < 0.01
    assert agree > 0.6


def test_float_input_and_small_image(hgmod):
    """(B,3,H,W) float32 input path, smallest supported map sizes (64 x 64 input -> 1 x 1 at the bottom)."""
    err, agree, gap, cerr = _compare(hgmod, 2, 64, 64, 5, seed=2, float_input=True)
    assert err < 0.01 and gap < 0.01


def test_batch_larger_than_chunk_is_consistent(hgmod, monkeypatch):
    """Images are independent: results do not depend on how the batch is cut into chunks
    (ragged last chunk, partially filled multi-image M tiles on the 8x8 / 4x4 / 2x2 levels)."""
    model = ohg.make_model(2, seed=3)
    img = ohg.to_uint8(ohg.synthetic_images(11, 128, 128, seed=4)).cuda()
    eng_a = hgmod.HourglassEngine(model.state_dict(), 128, 128, max_batch=11)
    ia, ca = eng_a.forward(img)
    monkeypatch.setenv("DF3D_HG_CHUNK", "4")                                   # 4 + 4 + 3
    eng_b = hgmod.HourglassEngine(model.state_dict(), 128, 128, max_batch=11)
    ib, cb = eng_b.forward(img)
    torch.cuda.synchronize()
    assert torch.equal(ia, ib) and torch.equal(ca, cb)
    assert eng_b.launches(11) == 3 * eng_a.launches(11)
    # two concurrent lanes (own stream, half of the SMs each), ragged split 6 + 5 and 2+2 / 2+2 / 2+1
    for chunk in ("8", "2"):
        monkeypatch.setenv("DF3D_HG_LANES", "2")
        monkeypatch.setenv("DF3D_HG_CHUNK", chunk)
        eng_c = hgmod.HourglassEngine(model.state_dict(), 128, 128, max_batch=11)
        ic, cc = eng_c.forward(img)
        torch.cuda.synchronize()
        assert torch.equal(ia, ic) and torch.equal(ca, cc)


@pytest.mark.parametrize("shape", [(2, 128, 128, 5), (2, 256, 512, 3), (8, 256, 256, 3), (2, 64, 64, 9)])
def test_fused_plans_are_bit_identical(hgmod, monkeypatch, shape):
    """The conv-chain plans (DF3D_HG_FUSE=1: point-wise chains behind stand-alone 3x3 convs, =2: 3x3-led
    chains, the default) keep every rounding point of the one-launch-per-conv plan (=0): the bf16
    intermediates that no longer go through HBM are rounded exactly as if they had been stored, so
    score maps, arg-max indices and confidences must be bit-identical."""
    stacks, H, W, B = shape
    model = ohg.make_model(stacks, seed=7)
    img = ohg.to_uint8(ohg.synthetic_images(B, H, W, seed=8)).cuda()
    flip = torch.tensor([i % 2 for i in range(B)], dtype=torch.uint8).cuda()
    outs = []
    # the one algebraic change of the default plan that moves a rounding point -- fc(res.conv3(t) + h) evaluated as
    # (W_fc W_3) t + W_fc h -- is switched off for the bit-identity part and compared separately below
    monkeypatch.setenv("DF3D_HG_NO_FC_MERGE", "1")
    for fuse in ("0", "1", "2"):
        monkeypatch.setenv("DF3D_HG_FUSE", fuse)
        eng = hgmod.HourglassEngine(model.state_dict(), H, W, max_batch=B)
        idx, conf, heat = eng.forward(img, flip=flip, return_heatmap=True)
        torch.cuda.synchronize()
        outs.append((idx.clone(), conf.clone(), heat[..., :19].clone(), eng.launches(B)))
        eng.close()
    for k in (1, 2):
        assert torch.equal(outs[0][2], outs[k][2]), f"score maps differ between DF3D_HG_FUSE=0 and {k}"
        assert torch.equal(outs[0][0], outs[k][0]) and torch.equal(outs[0][1], outs[k][1])
    assert outs[2][3] < outs[1][3] < outs[0][3]        # fewer launches the more is chained
    monkeypatch.delenv("DF3D_HG_NO_FC_MERGE")
    eng = hgmod.HourglassEngine(model.state_dict(), H, W, max_batch=B)
    idx, conf, heat = eng.forward(img, flip=flip, return_heatmap=True)
    torch.cuda.synchronize()
    rng_ = float(outs[0][2].max() - outs[0][2].min())
    err = float((heat[..., :19] - outs[0][2]).abs().max()) / rng_
    print(f"  default plan (fc . conv3 merged) vs the bit-identical plans: heat err {err:.4f} of range")
    assert err < 0.01
    eng.close()


def test_front_section_fusion_matches_unfused(hgmod, monkeypatch):
    """Default plan: layer1's conv3 + projection shortcut as one K-concatenated GEMM whose epilogue max-pools the tile
    (csrc/conv_gemm.cu: kb_split, pool2).  DF3D_HG_NO_FRONT_FUSE=1 runs the same layers as separate launches with the
    shortcut rounded to bf16 on its own: the two must agree to bf16 rounding noise, and the arg-max of their own
    maps must decode bit-exactly."""
    model = ohg.make_model(2, seed=11)
    img = ohg.to_uint8(ohg.synthetic_images(5, 256, 512, seed=12)).cuda()
    flip = torch.tensor([0, 1, 0, 1, 1], dtype=torch.uint8).cuda()
    outs = []
    for knob in (None, "1"):
        if knob:
            monkeypatch.setenv("DF3D_HG_NO_FRONT_FUSE", knob)
        eng = hgmod.HourglassEngine(model.state_dict(), 256, 512, max_batch=5)
        idx, conf, heat = eng.forward(img, flip=flip, return_heatmap=True)
        torch.cuda.synchronize()
        outs.append((idx.cpu().numpy(), heat[..., :19].cpu(), eng.launches(5)))
        eng.close()
    rng_ = float(outs[1][1].max() - outs[1][1].min())
    err = float((outs[0][1] - outs[1][1]).abs().max()) / rng_
    print(f"  front fusion vs separate launches: heat err {err:.4f} of range, launches {outs[0][2]} vs {outs[1][2]}")
    assert err < 0.01
    assert outs[0][2] == outs[1][2] - 2
    own_idx, _ = oargmax.heatmap_argmax(outs[0][1].permute(0, 3, 1, 2).contiguous().numpy())
    assert np.array_equal(outs[0][0], own_idx)


def _blob_batch(n, size, gen, device):
    """Joint k is a Gaussian blob somewhere inside cell k of a 5 x 4 grid; target = its heat-map at 1/4 resolution."""
    cells_x, cells_y = 5, 4
    cw, ch = size / cells_x, size / cells_y
    k = torch.arange(19, device=device)
    cx = ((k % cells_x).float() + 0.15 + 0.7 * torch.rand((n, 19), generator=gen, device=device)) * cw
    cy = ((k // cells_x).float() + 0.15 + 0.7 * torch.rand((n, 19), generator=gen, device=device)) * ch
    ys = torch.arange(size, dtype=torch.float32, device=device).view(1, 1, size, 1)
    xs = torch.arange(size, dtype=torch.float32, device=device).view(1, 1, 1, size)
    img = torch.exp(-((ys - cy.view(n, 19, 1, 1)) ** 2 + (xs - cx.view(n, 19, 1, 1)) ** 2) / (2 * 3.0 ** 2)).sum(1)
    img = (img + 0.03 * torch.randn((n, size, size), generator=gen, device=device)).clamp_(0, 1)
    hs = size // 4
    yh = torch.arange(hs, dtype=torch.float32, device=device).view(1, 1, hs, 1)
    xh = torch.arange(hs, dtype=torch.float32, device=device).view(1, 1, 1, hs)
    tgt = torch.exp(-((yh - (cy / 4).view(n, 19, 1, 1)) ** 2 + (xh - (cx / 4).view(n, 19, 1, 1)) ** 2) / (2 * 1.0 ** 2))
    return (img * 255).round().to(torch.uint8), tgt, torch.stack([cy / 4, cx / 4], dim=-1)


def test_trained_network_finds_the_joints(hgmod):
    """Seeded random weights give noise-like maps whose arg-max is a coin toss between near-ties.  Here the oracle
    network is first TRAINED (torch autograd on the GPU, test infrastructure only) to localise 19 blobs, so its maps
    are peaked like a real pose network's; then the CUDA engine (bf16 tensor-core path) must put the arg-max where
    the fp32 oracle puts it -- the reference's own tolerance is +-1 heat-map cell (tests/test_df3d.py:171) -- and
    where the blobs are."""
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    torch.manual_seed(0)
    model = ohg.make_model(2, seed=5).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    size = 128
    for step in range(400):
        img, tgt, _ = _blob_batch(16, size, gen, dev)
        x = (img.float() / 255.0 - 0.5).unsqueeze(1).expand(-1, 3, -1, -1)
        loss = sum(((o - tgt) ** 2).mean() for o in model(x))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    model.eval().cpu()
    img, tgt, pos = _blob_batch(24, size, gen, dev)
    eng = hgmod.HourglassEngine(model.state_dict(), size, size, max_batch=24)
    idx, conf, heat = eng.forward(img, return_heatmap=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        x = ohg.preprocess_u8(img.cpu())
        ref32 = model(x)[-1]
        ref16 = model(x, emulate_bf16=True)[-1]
    hs = size // 4
    got = idx.cpu().numpy()
    i32, c32 = oargmax.heatmap_argmax(ref32.numpy())
    i16, _ = oargmax.heatmap_argmax(ref16.numpy())
    rc = lambda a: np.stack(np.divmod(a, hs), axis=-1).astype(np.float64)
    d32 = np.abs(rc(got) - rc(i32)).max(axis=-1)
    d16 = np.abs(rc(got) - rc(i16)).max(axis=-1)
    dgt = np.abs(rc(got) - pos.cpu().numpy()).max(axis=-1)
    peaked = c32 > 0.3                                   # joints the trained oracle itself localises with a clear peak
    print(f"  trained 2-stack network: final loss {float(loss):.4f}, peaked joints {peaked.mean():.3f}; CUDA arg-max == fp32 oracle "
          f"{(d32 == 0)[peaked].mean():.3f}, within 1 cell {(d32 <= 1)[peaked].mean():.3f}; == bf16 oracle {(d16 == 0)[peaked].mean():.3f}; "
          f"within 1.5 cells of the blob {(dgt <= 1.5)[peaked].mean():.3f}")
    assert peaked.mean() > 0.8, "training did not produce peaked maps"
    assert (d32 <= 1)[peaked].mean() >= 0.99             # reference tolerance: atol 0.02 = +-1 cell of a 64-row map
    assert (d32 == 0)[peaked].mean() >= 0.85            # (the training run itself is not bit-reproducible: margins)
    assert (d16 == 0)[peaked].mean() >= 0.85 and (d16 <= 1)[peaked].mean() >= 0.99
    assert (dgt <= 1.5)[peaked].mean() >= 0.95
    # peak values: bf16 activations through ~100 convolutions move the (sharp) peaks by a few percent of the map's range
    # -- the reference's fp32-vs-fp32 tolerance on the confidence (atol 0.002, tests/test_df3d.py:177) is out of reach of
    # any bf16 network; against the oracle that rounds where the kernels round the bound is tight
    rngv = float(ref32.max() - ref32.min())
    _, c16 = oargmax.heatmap_argmax(ref16.numpy())
    e32 = float(np.abs(conf.cpu().numpy() - c32)[peaked].max()) / rngv
    e16 = float(np.abs(conf.cpu().numpy() - c16)[peaked].max()) / rngv
    print(f"  confidence: {e32:.4f} of range from the fp32 oracle, {e16:.4f} from the bf16-emulating oracle")
    assert e32 < 0.06 and e16 < 0.04                     # a trained (high-gain) network amplifies every bf16 rounding flip
    eng.close()
