"""Layer-level parity of the tcgen05 implicit-GEMM convolution against a float32 torch reference
on identical bf16 inputs/weights: the only difference is the fp32 summation order, so the bf16
outputs may differ by at most one rounding step."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hgmod(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from deepfly3d_b200 import hourglass

    return hourglass


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


CASES = [
    # B, H, W, Cin, Cout, k
    (2, 64, 64, 256, 128, 1),
    (2, 64, 64, 128, 128, 3),
    (1, 64, 64, 128, 256, 1),
    (2, 64, 64, 256, 256, 1),
    (3, 32, 32, 128, 128, 3),
    (5, 16, 16, 128, 128, 3),
    (5, 8, 8, 128, 128, 3),      # 2 images per M tile, odd batch -> partial tile
    (11, 4, 4, 128, 128, 3),     # 8 images per M tile
    (3, 4, 4, 256, 128, 1),
    (1, 128, 128, 64, 64, 3),
    (1, 128, 128, 64, 128, 1),
    (2, 64, 64, 256, 64, 1),     # intermediate score head (N = 64)
    (2, 64, 64, 64, 256, 1),     # score re-injection
    (1, 64, 128, 128, 128, 3),   # non-square map (reference heat-map 64 x 128)
    (1, 128, 128, 192, 64, 1),   # stem GEMM shape (K = 192)
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", CASES)
def test_conv_layer(hgmod, B, H, W, Cin, Cout, k):
    gen = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + k)
    x = _bf(torch.randn((B, H, W, Cin), generator=gen))
    w = torch.randn((Cout, Cin, k, k), generator=gen) * (1.0 / (Cin * k * k)) ** 0.5
    s1 = 1.0 + 0.1 * torch.randn(Cout, generator=gen)
    h1 = 0.1 * torch.randn(Cout, generator=gen)
    s2 = 1.0 + 0.1 * torch.randn(Cout, generator=gen)
    h2 = 0.1 * torch.randn(Cout, generator=gen)
    res = _bf(torch.randn((B, H, W, Cout), generator=gen))

    out, act = hgmod.conv2d_nhwc_bf16(x.cuda().to(torch.bfloat16), w, s1, h1, relu1=False,
                                      residual=res.cuda().to(torch.bfloat16), scale2=s2, shift2=h2)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2), _bf(w), padding=k // 2).permute(0, 2, 3, 1)
    ref = ref * s1 + h1 + res
    ref_act = torch.relu(_bf(ref) * s2 + h2)
    got, got_act = out.float().cpu(), act.float().cpu()
    # one bf16 ulp (2^-8 relative) + fp32 summation noise
    tol = 2.0 ** -7
    err = ((got - ref).abs() / (ref.abs() + 1.0)).max().item()
    err_act = ((got_act - ref_act).abs() / (ref_act.abs() + 1.0)).max().item()
    assert err < tol, f"raw output mismatch {err}"
    assert err_act < tol, f"activated output mismatch {err_act}"
    # the overwhelming majority must be bit-identical after rounding
    same = (got == _bf(ref)).float().mean().item()
    assert same > 0.98, f"only {same:.4f} of the outputs are bit-identical"


def test_conv_relu_no_residual(hgmod):
    gen = torch.Generator().manual_seed(7)
    x = _bf(torch.randn((2, 16, 16, 128), generator=gen))
    w = torch.randn((128, 128, 3, 3), generator=gen) * 0.03
    s1 = torch.ones(128)
    h1 = 0.05 * torch.randn(128, generator=gen)
    out, act = hgmod.conv2d_nhwc_bf16(x.cuda().to(torch.bfloat16), w, s1, h1, relu1=True)
    assert act is None
    ref = torch.relu(F.conv2d(x.permute(0, 3, 1, 2), _bf(w), padding=1).permute(0, 2, 3, 1) + h1)
    assert ((out.float().cpu() - ref).abs() / (ref.abs() + 1.0)).max().item() < 2.0 ** -7


def test_conv_rejects_unsupported(hgmod):
    from deepfly3d_b200._lib import Df3dError

    x = torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(Df3dError):
        hgmod.conv2d_nhwc_bf16(x, torch.zeros((64, 48, 1, 1)), torch.ones(64), torch.zeros(64))
