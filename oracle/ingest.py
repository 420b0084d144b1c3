"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the image resize of the ingest step.

The reference reads camera_C_img_I.jpg and hands the images to df2d, which resizes them to the network input
(call site df3d/core.py:177-185; the resize mode inside the un-vendored df2d is not verifiable offline, SURVEY.md
section 8-a2).  This port's definition is what deepfly3d_b200/inference.py has always done on the host:
cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR) on uint8 gray images.  `resize_bilinear_u8` restates
OpenCV's fixed-point algorithm for that case (modules/imgproc/src/resize.cpp: 11-bit coefficients, rows
accumulated in int32, the 22-bit vertical cast) so that the CUDA kernel can be checked bit for bit, and is itself
pinned against cv2.resize in tests/test_oracle_ingest.py.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _coefficients(src, dst, clamp_fraction):
    """-> (i0[dst], i1[dst], a0[dst], a1[dst]) int32: the two source indices and 11-bit weights of OpenCV's
    INTER_LINEAR for one axis.  Columns (clamp_fraction=True): a source position left of pixel 0 / right of the
    last pixel snaps to that pixel with weight 2048.  Rows (False): the fraction is kept and the two row indices
    are clamped separately, so a border row is blended with itself -- two truncated terms, not one."""
    scale = float(src) / float(dst)
    i0 = np.empty(dst, np.int32)
    i1 = np.empty(dst, np.int32)
    a0 = np.empty(dst, np.int32)
    a1 = np.empty(dst, np.int32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_fraction:
            if s < 0:
                s, f = 0, np.float32(0.0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0.0)
        i0[d] = min(max(s, 0), src - 1)
        i1[d] = min(max(s + 1, 0), src - 1)
        # saturate_cast<short>(float) = round half to even
        a0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(COEF_SCALE)))
        a1[d] = int(np.rint(f * np.float32(COEF_SCALE)))
    return i0, i1, a0, a1


def resize_bilinear_u8(img, out_hw):
    """img (H, W) uint8 -> (Hd, Wd) uint8, bit-identical to cv2.resize(img, (Wd, Hd), interpolation=INTER_LINEAR)."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    Hs, Ws = img.shape
    Hd, Wd = out_hw
    if (Hs, Ws) == (Hd, Wd):
        return img.copy()
    x0, x1, xa0, xa1 = _coefficients(Ws, Wd, True)
    y0, y1, ya0, ya1 = _coefficients(Hs, Hd, False)
    src = img.astype(np.int32)
    rows = src[:, x0] * xa0[None, :] + src[:, x1] * xa1[None, :]           # (Hs, Wd), scale 2^11
    r0 = rows[y0] >> 4
    r1 = rows[y1] >> 4
    out = (((ya0[:, None] * r0) >> 16) + ((ya1[:, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
