"""Pins oracle/ingest.py (the CPU restatement of the loader's resize) against cv2.resize itself -- the
definition the host loader has always used (deepfly3d_b200/inference.py; reference call site
df3d/core.py:177-185) -- on random images and on the reference's own sample frames.  Bit-exact."""
import glob
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle.ingest import resize_bilinear_u8

HERE = os.path.dirname(os.path.abspath(__file__))

SHAPES = [
    (480, 960, 256, 512),   # the reference's frames -> network input (config 1)
    (480, 960, 256, 256),
    (128, 128, 256, 256),   # up-scaling: border rows blend a row with itself
    (100, 100, 256, 256),
    (77, 131, 64, 64),
    (512, 512, 256, 256),   # exact 2:1
    (33, 47, 256, 512),
    (480, 960, 480, 960),   # identity
]


@pytest.mark.parametrize("shape", SHAPES)
def test_resize_oracle_matches_cv2_bit_exact(shape):
    hs, ws, hd, wd = shape
    rng = np.random.default_rng(hs * 1000 + wd)
    img = rng.integers(0, 256, (hs, ws), dtype=np.uint8)
    ref = cv2.resize(img, (wd, hd), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(resize_bilinear_u8(img, (hd, wd)), ref)


def test_resize_oracle_on_reference_frames():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "images", "*.jpg")))[:7]
    assert files, "image fixtures missing"
    for f in files:
        img = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        ref = cv2.resize(img, (512, 256), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(resize_bilinear_u8(img, (256, 512)), ref)


def test_resize_oracle_extremes():
    for v in (0, 255):
        img = np.full((37, 53), v, np.uint8)
        assert np.array_equal(resize_bilinear_u8(img, (64, 128)), np.full((64, 128), v, np.uint8))
