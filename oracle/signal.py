"""Oracle restatement of the reference's post-processing filters.

Test infrastructure (see ``oracle/__init__.py``).  **Parity pinned**: ``tests/golden/signal.npz`` holds outputs of
the reference's own ``df3d/signal_util.py`` / ``df3d/plot_util.py`` run in the build container
(``tests/golden/make_golden_signal.py``); ``tests/test_oracle_signal.py`` checks this restatement against them.

  one_euro_batch     df3d/signal_util.py:31-66 (OneEuroFilter), :69-100 (filter_batch: timestamps (i+1)*0.1),
                     :103-132 (filter_batch_2d: timestamps i*0.1) -- including the reference's quirk that the
                     sampling frequency is re-derived from the time stamps (1 / 0.1 = 10 Hz from the second
                     sample on, whatever `freq` says) and that a time stamp of 0.0 is "falsy" and skips that update
  smooth_pose2d      df3d/signal_util.py:135-160 (windowed std -> Gaussian sigma 7 or 0.1, scipy
                     gaussian_filter1d(mode='nearest') of the 20-sample window, centre tap)
  normalize_pose_3d  df3d/plot_util.py:85-91, 10-17 (median-centre, then y <- -z, z <- -y)
"""
import math

import numpy as np


def _alpha(freq, cutoff):
    te = 1.0 / freq
    tau = 1.0 / (2 * math.pi * cutoff)
    return 1.0 / (1.0 + tau / te)


def one_euro_batch(pts, freq=100.0, mincutoff=0.1, beta=2.0, dcutoff=1.0, t_first=1):
    """pts (T, J, D) -> filtered (T, J, D).  t_first = 1: filter_batch ((i+1)*0.1); 0: filter_batch_2d (i*0.1)."""
    pts = np.asarray(pts, dtype=np.float64)
    T = pts.shape[0]
    out = np.zeros_like(pts)
    f = np.full(pts.shape[1:], float(freq))
    x_y = x_s = dx_s = None
    last = None
    for i in range(T):
        ts = (i + t_first) * 0.1
        if last and ts:                       # python truthiness: None and 0.0 skip the update
            f = np.full(pts.shape[1:], 1.0 / (ts - last))
        last = ts
        x = pts[i]
        dx = np.zeros_like(x) if x_y is None else (x - x_y) * f
        a_d = _alpha(f, dcutoff)
        edx = dx if dx_s is None else a_d * dx + (1.0 - a_d) * dx_s
        dx_s = edx
        cutoff = mincutoff + beta * np.abs(edx)
        a = 1.0 / (1.0 + (1.0 / (2 * math.pi * cutoff)) / (1.0 / f))
        s = x if x_s is None else a * x + (1.0 - a) * x_s
        x_y, x_s = x, s
        out[i] = s
    return out


def gaussian_weights(sigma, truncate=4.0):
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return w / w.sum(), radius


def smooth_pose2d(points2d, window_size=20, pad=20, std_thr=5):
    p = np.asarray(points2d, dtype=np.float64)
    T = p.shape[0]
    padded = np.concatenate([np.repeat(p[:1], pad, 0), p, np.repeat(p[-1:], pad, 0)], axis=0)
    out = p.copy()
    half = window_size // 2
    w7, r7 = gaussian_weights(7.0)
    for t in range(T):
        seg = padded[t + pad - half: t + pad + half]                 # (window, J, 2)
        std = seg.std(axis=0)
        idx = np.clip(half + np.arange(-r7, r7 + 1), 0, window_size - 1)
        smooth = np.tensordot(w7, seg[idx], axes=(0, 0))
        out[t] = np.where(std < std_thr, smooth, seg[half])          # sigma 0.1 -> radius 0 -> the centre tap itself
    return out


def normalize_pose_3d(points3d, rotate=True):
    p = np.asarray(points3d, dtype=np.float64).copy()
    p -= np.median(p.reshape(-1, 3), axis=0)
    if rotate:
        y, z = p[..., 1].copy(), p[..., 2].copy()
        p[..., 1] = -z
        p[..., 2] = -y
    return p
