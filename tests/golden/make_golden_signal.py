"""Golden vectors for the post-processing rows, produced by RUNNING THE REFERENCE'S OWN
CODE: df3d.signal_util (One-Euro filters, smooth_pose2d), df3d.plot_util.normalize_pose_3d and
df3d.procrustes.procrustes_seperate import in the build container with PYTHONPATH=/root/reference (they are
pure numpy / scipy).  Run in the build container only:

    PYTHONPATH=/root/reference python tests/golden/make_golden_signal.py

Inputs are seeded; the file written is tests/golden/signal.npz:
  pts3d_in (T=48,38,3)       a random walk around the golden 3-D joints (points3d_wo_procrustes tiled)
  filter_batch_out           df3d/signal_util.py:69-100   filter_batch(pts3d_in)
  filter_batch_freq30_out    ... with freq=30
  get_points3d_out           df3d/core.py:332-343: procrustes_seperate -> normalize_pose_3d(rotate=True) -> filter_batch
  normalize_rotate_out       df3d/plot_util.py:85-91 with rotate=True on pts3d_in
  pts2d_in (T=64,38,2)       pixel tracks: smooth random walk + a few jumps (both branches of the std threshold)
  smooth_pose2d_out          df3d/signal_util.py:135-160
  filter_batch_2d_out        df3d/signal_util.py:103-132
"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    from df3d import plot_util, procrustes, signal_util

    rng = np.random.default_rng(7)
    with open(f"{REF}/tests/data/reference_df3d/df3d_result_3d.pkl", "rb") as f:
        r3 = pickle.load(f)
    base = np.asarray(r3["points3d_wo_procrustes"], dtype=np.float64)           # (15,38,3)
    T = 48
    walk = np.cumsum(rng.normal(scale=0.01, size=(T, 38, 3)), axis=0)
    pts3d = base[np.arange(T) % base.shape[0]] + walk
    out = {"pts3d_in": pts3d}
    out["filter_batch_out"] = signal_util.filter_batch(pts3d.copy())
    out["filter_batch_freq30_out"] = signal_util.filter_batch(pts3d.copy(), freq=30)
    out["normalize_rotate_out"] = plot_util.normalize_pose_3d(pts3d.copy(), rotate=True)
    p = procrustes.procrustes_seperate(pts3d.copy())
    out["procrustes_out"] = p.copy()
    p = plot_util.normalize_pose_3d(p, rotate=True)
    out["get_points3d_out"] = signal_util.filter_batch(p)

    T2 = 64
    pts2d = np.cumsum(rng.normal(scale=0.6, size=(T2, 38, 2)), axis=0) + rng.uniform(50, 900, size=(1, 38, 2))
    jumps = rng.random((T2, 38)) < 0.04
    pts2d[jumps] += rng.normal(scale=60.0, size=(int(jumps.sum()), 2))
    out["pts2d_in"] = pts2d
    out["smooth_pose2d_out"] = signal_util.smooth_pose2d(pts2d.copy())
    out["filter_batch_2d_out"] = signal_util.filter_batch_2d(pts2d.copy())
    np.savez_compressed(os.path.join(HERE, "signal.npz"), **out)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
