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


def test_video_reader_streams_frames_in_order(tmp_path):
    """camera_N.mp4 -> frame blocks (config 5 ingest): camera and frame order, gray conversion, block boundaries."""
    import torch

    from deepfly3d_b200.inference import VideoReader

    T, H, W = 9, 64, 96
    for c in range(7):
        vw = cv2.VideoWriter(str(tmp_path / f"camera_{c}.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (W, H))
        if not vw.isOpened():
            pytest.skip("no mp4 encoder in this OpenCV build")
        for t in range(T):
            vw.write(np.full((H, W, 3), 16 + 32 * c + 3 * t, dtype=np.uint8))
        vw.release()
    with VideoReader(str(tmp_path)) as vr:
        assert vr.num_frames == T and vr.shape == (H, W)
        buf = torch.zeros((7, 4, H, W), dtype=torch.uint8)
        seen = np.zeros((7, T))
        for t0, t1 in [(0, 4), (4, 8), (8, 9)]:
            vr.wait(vr.read_block_async(t0, t1, buf))
            seen[:, t0:t1] = buf[:, :t1 - t0].float().mean(dim=(2, 3)).numpy()
        with pytest.raises(ValueError):
            vr.wait(vr.read_block_async(3, 4, buf))                        # out of order
    expect = 16 + 32 * np.arange(7)[:, None] + 3 * np.arange(T)[None, :]
    assert np.abs(seen - expect).max() <= 3.5                              # lossy codec (limited-range YUV), flat frames
    assert np.all(np.diff(seen, axis=1) > 0) and np.all(np.diff(seen, axis=0) > 0)   # frame and camera order


def test_block_plan_of_the_streaming_loader():
    """Host logic of inference_folder: engine block from the input size (1 792 images at 256 x 256, scaled by area, never
    below the reference's batch_size), quarter-size blocks for the host decode, blocks covering [0, T) in order."""
    from deepfly3d_b200 import inference

    assert inference.block_frames_for(256, 256, 1000) == 256
    assert inference.block_frames_for(256, 512, 1000) == 128
    assert inference.block_frames_for(256, 512, 3) == 3                      # never more than the recording
    assert inference.block_frames_for(2048, 2048, 1000, batch_size=8) == 4   # 28 images: 4 whole frames
    assert inference.block_frames_for(2048, 2048, 1000, batch_size=64) == 9  # a larger batch_size raises the block
    assert inference.stream_block_frames(128) == 32 and inference.stream_block_frames(128, gpu_decode=True) == 128
    assert inference.stream_block_frames(3) == 3
    assert inference.plan_blocks(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert inference.plan_blocks(0, 4) == []
