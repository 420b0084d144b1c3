"""Regenerate the committed fixtures from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``

Sources (all under /root/reference):
  data/calib.pkl                                   -> calib.npz         (BA initial state, core.py:234-242)
  data/df3d_result.pkl['points3d']                 -> template.npz      (procrustes template, procrustes.py:38-48)
  tests/data/reference_df3d/df3d_result_2d.pkl     -> result_2d.npz     (input of test_calibration, test_df3d.py:213-214)
  tests/data/reference_df3d/df3d_result_3d.pkl     -> result_3d.npz     (expected of test_calibration, test_df3d.py:221-243)

  tests/data/reference/camera_{0..6}_img_{0..2}.jpg -> images/          (21 of the 105 sample JPEGs, plumbing input
                                                                         of test_pose_estimation, test_df3d.py:150-160)

The pickles are converted to plain ``.npz`` so that no reference class is needed
to load them.  ``calib.npz`` and ``template.npz`` are also copied into
``deepfly3d_b200/data/`` because the product path needs them at run time exactly
as the reference reads ``data/*.pkl``.
"""
import os
import pickle
import shutil

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DATA = os.path.join(HERE, "..", "..", "deepfly3d_b200", "data")


def cams_to_arrays(d):
    out = {}
    for key in ("R", "tvec", "intr", "distort"):
        out[key] = np.stack([np.asarray(d[c][key], dtype=np.float64) for c in range(7)])
    return out


def main():
    with open(f"{REF}/data/calib.pkl", "rb") as f:
        calib = pickle.load(f)
    np.savez(os.path.join(HERE, "calib.npz"), **cams_to_arrays(calib))

    with open(f"{REF}/data/df3d_result.pkl", "rb") as f:
        tmpl = pickle.load(f)
    np.savez(os.path.join(HERE, "template.npz"), points3d=np.asarray(tmpl["points3d"], dtype=np.float64))

    with open(f"{REF}/tests/data/reference_df3d/df3d_result_2d.pkl", "rb") as f:
        r2 = pickle.load(f)
    np.savez(
        os.path.join(HERE, "result_2d.npz"),
        points2d=r2["points2d"],
        camera_ordering=r2["camera_ordering"],
        heatmap_confidence=r2["heatmap_confidence"],
    )

    with open(f"{REF}/tests/data/reference_df3d/df3d_result_3d.pkl", "rb") as f:
        r3 = pickle.load(f)
    np.savez(
        os.path.join(HERE, "result_3d.npz"),
        points3d=r3["points3d"],
        points3d_wo_procrustes=r3["points3d_wo_procrustes"],
        points2d=r3["points2d"],
        camera_ordering=r3["camera_ordering"],
        heatmap_confidence=r3["heatmap_confidence"],
        **cams_to_arrays(r3),
    )

    img_dir = os.path.join(HERE, "images")
    os.makedirs(img_dir, exist_ok=True)
    for cam in range(7):
        for img_id in range(3):
            name = f"camera_{cam}_img_{img_id}.jpg"
            shutil.copy(f"{REF}/tests/data/reference/{name}", os.path.join(img_dir, name))
            os.chmod(os.path.join(img_dir, name), 0o644)

    os.makedirs(PKG_DATA, exist_ok=True)
    for name in ("calib.npz", "template.npz"):
        shutil.copy(os.path.join(HERE, name), os.path.join(PKG_DATA, name))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
