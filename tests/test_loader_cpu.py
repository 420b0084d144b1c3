"""Host side of the image loader (deepfly3d_b200/inference.py): file naming and the native-size read that feeds
the device-side resize.  Mirrors what df2d's loader is given by df3d/core.py:177-185 (camera_C_img_I.jpg)."""
import os
import shutil

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = os.path.join(HERE, "golden", "images")


def test_read_images_native_size(tmp_path):
    from deepfly3d_b200.inference import image_name, read_images

    for f in os.listdir(IMAGES):
        shutil.copy(os.path.join(IMAGES, f), tmp_path / f)
    out = read_images(str(tmp_path), max_img_id=2, pin_memory=False)
    assert tuple(out.shape) == (7, 3, 480, 960) and out.dtype.is_floating_point is False
    ref = cv2.imread(image_name(str(tmp_path), 4, 1), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(out[4, 1].numpy(), ref)


def test_read_images_missing_file(tmp_path):
    from deepfly3d_b200.inference import read_images

    with pytest.raises(FileNotFoundError):
        read_images(str(tmp_path), max_img_id=0, pin_memory=False)


def test_image_name_accepts_zero_padded_ids(tmp_path):
    from deepfly3d_b200.inference import image_name

    (tmp_path / "camera_2_img_000007.jpg").write_bytes(b"x")
    assert image_name(str(tmp_path), 2, 7).endswith("camera_2_img_000007.jpg")
    (tmp_path / "camera_2_img_7.jpg").write_bytes(b"x")
    assert image_name(str(tmp_path), 2, 7).endswith("camera_2_img_7.jpg")
