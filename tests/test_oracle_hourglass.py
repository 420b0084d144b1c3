"""Oracle hourglass: architecture bookkeeping (FLOPs / parameter counts of SURVEY.md section 8(d)),
decode rule, and consistency of the bf16-emulating mode."""
import numpy as np
import torch

from oracle import argmax as oargmax
from oracle import hourglass as ohg


def test_flops_and_params_match_survey():
    m8 = ohg.make_model(8)
    assert abs(ohg.conv_flops(m8, 256, 256) / 1e9 - 54.975) < 1e-3
    assert abs(sum(p.numel() for p in m8.parameters()) / 1e6 - 25.446) < 1e-3
    m2 = ohg.make_model(2)
    assert abs(ohg.conv_flops(m2, 256, 512) / 1e9 - 33.376) < 1e-3
    assert abs(sum(p.numel() for p in m2.parameters()) / 1e6 - 6.573) < 1e-3
    n_conv = sum(1 for m in m8.modules() if isinstance(m, torch.nn.Conv2d))
    n_3x3 = sum(1 for m in m8.modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3))
    assert (n_conv, n_3x3) == (378, 115)


def test_output_shape_and_bf16_mode_close():
    m = ohg.make_model(2)
    x = ohg.preprocess_u8(ohg.to_uint8(ohg.synthetic_images(1, 128, 128, seed=3)))
    with torch.no_grad():
        a = m(x)
        b = m(x, emulate_bf16=True)
    assert len(a) == 2 and a[-1].shape == (1, 19, 32, 32)
    rel = (a[-1] - b[-1]).abs().max() / a[-1].abs().max()
    assert rel < 0.1


def test_argmax_first_occurrence():
    hm = np.zeros((1, 2, 4, 8), dtype=np.float32)
    hm[0, 0, 1, 3] = 2.0
    hm[0, 0, 2, 5] = 2.0          # tie -> first (row-major) wins
    hm[0, 1, 3, 7] = -1.0
    hm[0, 1] -= 5.0
    idx, conf = oargmax.heatmap_argmax(hm)
    assert idx[0, 0] == 1 * 8 + 3 and conf[0, 0] == 2.0
    assert idx[0, 1] == 0 and conf[0, 1] == -5.0
