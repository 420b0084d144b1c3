"""Host logic of the Core mirror that needs no GPU (reference df3d/core.py:65-126, 325-330, 461-475 and
df3d/os_util.py:7-23): folder discovery, image count and shape, default camera ordering, result file name,
resume from a result pickle, --delete-images."""
import os
import pickle
import shutil

import numpy as np
import pytest

pytest.importorskip("cv2")

HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = os.path.join(HERE, "golden", "images")


@pytest.fixture()
def working(tmp_path):
    d = tmp_path / "sample" / "test"
    d.mkdir(parents=True)
    for f in os.listdir(IMAGES):
        shutil.copy(os.path.join(IMAGES, f), d / f)
    return str(d)


def test_core_discovers_the_folder(working):
    from deepfly3d_b200.core import Core

    core = Core(input_folder=working, output_folder=working + "_out", num_images_max=0, camera_ordering=None)
    assert core.num_images == 3 and core.max_img_id == 2          # binary search over the file names (os_util.py:7-23)
    assert core.image_shape == [960, 480]                          # [W, H] (core.py:91-97)
    assert np.all(core.camera_ordering == np.arange(7))
    assert core.save_path.endswith("df3d_result_{}.pkl".format(working.replace("/", "_")))   # core.py:325-330
    assert core.camNet is None and core.points2d is None           # nothing to resume from


def test_core_num_images_max_and_missing_images(working, tmp_path):
    from deepfly3d_b200.core import Core

    core = Core(input_folder=working, num_images_max=2)
    assert core.num_images == 2 and core.max_img_id == 1
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):                          # os_util.py:19-21 "No image found."
        Core(input_folder=str(empty))


def test_core_resumes_from_a_result_pickle(working, golden):
    """Core.__init__ reloads points2d / heatmap_confidence (/ points3d) from df3d_result*.pkl (core.py:108-126)."""
    from deepfly3d_b200.core import Core

    out = working + "_df3d"
    os.makedirs(out)
    r2 = golden["result_2d"]
    T = 3
    blob = {"points2d": r2["points2d"][:, :T], "heatmap_confidence": r2["heatmap_confidence"][:, :T],
            "camera_ordering": np.arange(7)}
    name = "df3d_result_{}.pkl".format(working.replace("/", "_"))
    with open(os.path.join(out, name), "wb") as f:
        pickle.dump(blob, f)
    core = Core(input_folder=working, output_folder=out)
    assert core.camNet is None                                     # a 2-D-only pickle carries no calibration (test_df3d.py:201)
    assert np.array_equal(core.points2d, blob["points2d"]) and np.array_equal(core.conf, blob["heatmap_confidence"])


def test_delete_images_keeps_other_files(working):
    from deepfly3d_b200.core import Core

    core = Core(input_folder=working)
    keep = os.path.join(working, "camera_0.mp4")
    open(keep, "wb").write(b"x")
    core.delete_images()                                           # core.py:461-475: only cameras that have a video
    left = os.listdir(working)
    assert "camera_0.mp4" in left
    assert not [f for f in left if f.startswith("camera_0_img_")]
    assert len([f for f in left if f.endswith(".jpg")]) == 6 * 3
