"""Drop-in mirror of ``df3d.core.Core`` (reference df3d/core.py:62-203, 229-250, 325-369) with the
df2d / pyba internals replaced by the CUDA path.

Same constructor, attributes (``points2d``, ``conf``, ``points3d``, ``camNet``,
``camera_ordering``, ``image_shape``, ``num_images``, ``max_img_id``, ``save_path``) and methods
(``pose2d_estimation``, ``calibrate_calc``, ``save``, ``get_points3d``), same result-pickle schema
(README.md:324-328).  GUI / plotting / manual-correction helpers are out of scope (SURVEY.md
section 2) and are not reproduced.
"""
import glob
import logging
import os
import pickle
import re
from typing import List, Optional

import numpy as np

from .skeleton import HEATMAP_SHAPE, NUM_CAMERAS, NUM_JOINTS

logger = logging.getLogger("df3d.logger")

_KNOWN_ORDERINGS = [          # df3d/core.py:33-41
    (r"/CLC/", [0, 6, 5, 4, 3, 2, 1]),
    (r"/FA/", [6, 5, 4, 3, 2, 1, 0]),
    (r"/SG/", [6, 5, 4, 3, 2, 1, 0]),
    (r"Laura", [0, 6, 5, 4, 3, 2, 1]),
    (r"AYMANNS_Florian", [6, 5, 4, 3, 2, 1, 0]),
    (r"sample/test", [0, 1, 2, 3, 4, 5, 6]),
    (r"/JB/", [6, 5, 4, 3, 2, 1, 0]),
]


def find_default_camera_ordering(input_folder):
    """Infers the camera ordering from the folder path (df3d/core.py:24-59)."""
    folder = str(input_folder)
    for regex, order in _KNOWN_ORDERINGS:
        if re.search(regex, folder):
            return np.array(order)
    raise NotImplementedError(
        f"Cannot find camera ordering for folder {input_folder}. Please set your camera ordering using the"
        " --order flag. Example usage is df3d-cli /your/path/images/ --order 0 1 2 3 4 5 6")


def _image_exists(path, img_id):
    return any(os.path.isfile(os.path.join(path, f"camera_{c}_img_{img_id}.jpg")) for c in range(NUM_CAMERAS)) or \
        os.path.isfile(os.path.join(path, f"camera_0_img_{img_id:06d}.jpg"))


def get_max_img_id(path):
    """Binary search over image ids (df3d/os_util.py:7-23)."""
    lo, hi = 0, 100000
    cur = (lo + hi) // 2
    while hi - lo > 1:
        if _image_exists(path, cur):
            lo = cur
        else:
            hi = cur
        cur = (lo + hi) // 2
    if not _image_exists(path, cur):
        raise FileNotFoundError("No image found.")
    return cur


def _read_image_shape(path):
    import cv2

    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    return None if img is None else [int(img.shape[1]), int(img.shape[0])]


class Core:
    """Main interface to the 2d and 3d pose estimation (reference df3d/core.py:62)."""

    def __init__(self, input_folder: str, output_folder: Optional[str] = None, num_images_max: Optional[int] = None,
                 camera_ordering: List[int] = [0, 1, 2, 3, 4, 5, 6], state_dict=None, weights=None, mean=None,
                 gpu_decode=False, stream_videos=False, block_frames=None, ba_solver="lsmr"):
        """Same four arguments as the reference.  Extra keywords (the reference reads them from its config,
        df3d/config.py:30-39): `weights` = hourglass checkpoint (sh8_deepfly.tar layout) or `state_dict`,
        `mean` = per-channel mean or the path of a mean.pth.tar, `gpu_decode` = nvJPEG instead of libjpeg (True: hardware JPEG engines if present; "hardware" / "default"),
        `block_frames` = frames per streamed block (default: sized for the GPU), `ba_solver` = "lsmr" (SciPy's own
        truncated step, the default) or "exact" (see CameraNetwork.bundle_adjust),
        `stream_videos` = decode camera_N.mp4 straight into the pipeline instead of expanding them to JPEG files
        first (opt-in: the reference's frames have been through ffmpeg's MJPEG encoder once more)."""
        self.input_folder = input_folder
        self.output_folder = self.input_folder + "_df3d" if output_folder is None else output_folder
        self._stream_videos = bool(stream_videos)
        if not self._stream_videos:
            self.expand_videos()
        self.fps = self.get_fps()
        self.num_images_max = num_images_max if num_images_max is not None else 0
        video_shape = None
        if self._stream_videos:
            from .inference import VideoReader

            with VideoReader(self.input_folder) as vr:
                if vr.num_frames < 1:
                    raise FileNotFoundError("No image found.")
                self.max_img_id, video_shape = vr.num_frames - 1, [vr.shape[1], vr.shape[0]]
        else:
            self.max_img_id = get_max_img_id(self.input_folder)
        if self.num_images_max > 0:
            self.num_images = min(self.num_images_max, self.max_img_id + 1)
            self.max_img_id = self.num_images - 1
        else:
            self.num_images = self.max_img_id + 1
        image_path = os.path.join(self.input_folder, "camera_{cam_id}_img_{img_id}.jpg")
        image0 = image_path.format(cam_id=0, img_id=0)
        shape = video_shape if video_shape is not None else (_read_image_shape(image0) if os.path.exists(image0) else None)
        if shape is None:
            raise ValueError(f"Image shape not specified and could not be read from {image0}")
        self.image_shape = shape                       # [W, H], e.g. [960, 480]
        self.camera_ordering = self.setup_camera_ordering(camera_ordering)
        self._state_dict, self._weights, self._mean, self._gpu_decode = state_dict, weights, mean, gpu_decode
        self._block_frames, self._ba_solver = block_frames, ba_solver
        self.ingest_stats = {}

        self.camNet = None
        self.points2d = None
        self.points3d = None
        self.conf = None
        if os.path.exists(self.save_path):            # resume from a previous run (core.py:108-126)
            with open(self.save_path, "rb") as f:
                res = pickle.load(f)
            self.points2d = res["points2d"]
            self.conf = res["heatmap_confidence"]
            if "points3d" in res:
                self.points3d = res["points3d"]
            if all(c in res for c in range(NUM_CAMERAS)):
                from .camera_network import CameraNetwork

                self.camNet = CameraNetwork(res["points2d"] * self.image_shape[::-1], calib=res, image_path=image_path)

    # ------------------------------------------------------------------ properties
    @property
    def input_folder(self):
        return self._input_folder

    @input_folder.setter
    def input_folder(self, value):
        value = os.path.abspath(value).rstrip("/")
        assert os.path.isdir(value), f"Not a directory {value}"
        self._input_folder = value

    @property
    def output_folder(self):
        return self._output_folder

    @output_folder.setter
    def output_folder(self, value):
        os.makedirs(value, exist_ok=True)
        value = os.path.abspath(value).rstrip("/")
        assert os.path.isdir(value), f"Not a directory {value}"
        self._output_folder = value

    @property
    def number_of_joints(self):
        return NUM_JOINTS

    @property
    def has_pose(self):
        return True

    @property
    def has_calibration(self):
        return self.camNet.has_calibration()

    @property
    def save_path(self):
        return os.path.join(self.output_folder, "df3d_result_{}.pkl".format(self.input_folder.replace("/", "_")))

    # ------------------------------------------------------------------ public methods
    def pose2d_estimation(self, batch_size: int = 8, disable_pin_memory: bool = False):
        """2-D pose on every image of the folder, then the 19 -> 38 packing (core.py:170-203)."""
        import torch

        from . import ops
        from .inference import inference_folder

        flip = [cam for index, cam in enumerate(self.camera_ordering) if index > 3]
        p19, conf = inference_folder(
            folder=self.input_folder, camera_ids_to_flip=flip, return_heatmap=False, return_confidence=True,
            max_img_id=self.max_img_id, batch_size=batch_size, disable_pin_memory=disable_pin_memory,
            state_dict=self._state_dict, weights=self._weights, mean=self._mean, gpu_decode=self._gpu_decode,
            block_frames=self._block_frames,
            stats=self.ingest_stats, source="videos" if self._stream_videos else "images")
        self.conf = conf
        # packing runs on the device from the integer arg-max indices (bit-exact with core.py:187-203)
        Hh, Wh = HEATMAP_SHAPE
        idx = np.round(p19[..., 0] * Hh).astype(np.int32) * Wh + np.round(p19[..., 1] * Wh).astype(np.int32)
        T = idx.shape[1]
        p2d, _ = ops.pack_points2d(torch.as_tensor(idx.reshape(NUM_CAMERAS * T, -1)).cuda(), NUM_CAMERAS, T,
                                   (Hh, Wh), self.camera_ordering, self.image_shape)
        self.points2d = p2d.cpu().numpy()

    def calibrate_calc(self, min_img_id, max_img_id):
        """Bundle adjustment from the packaged initial calibration (core.py:229-250; like the
        reference both arguments are ignored)."""
        from .camera_network import CameraNetwork
        from .pipeline import load_default_calib

        calib = load_default_calib()
        calib_reordered = {
            int(cidx): {k: calib[k][idx] for k in ("R", "tvec", "intr", "distort")}
            for idx, cidx in enumerate(self.camera_ordering)
        }
        image_path = os.path.join(self.input_folder, "camera_{cam_id}_img_{img_id}.jpg")
        self.camNet = CameraNetwork(self.points2d * self.image_shape[::-1], calib=calib_reordered, image_path=image_path)
        self.camNet.bundle_adjust(update_intrinsic=False, update_distort=False, solver=self._ba_solver)
        print(f"Reprojection error is {self.camNet.reprojection_error()}")

    def get_points3d(self):
        """(T,38,3) joints for display (core.py:332-343): procrustes registration, median-centring + axis swap
        (df3d/plot_util.py:85-91, 10-17: y <- -z, z <- -y), One-Euro filter (df3d/signal_util.py:69-100)."""
        import torch

        from . import ops

        pts = ops.procrustes(torch.as_tensor(np.ascontiguousarray(self.camNet.points3d, dtype=np.float64)).cuda())
        p = pts.cpu().numpy()
        p -= np.median(p.reshape(-1, 3), axis=0)
        y, z = p[..., 1].copy(), p[..., 2].copy()
        p[..., 1], p[..., 2] = -z, -y
        return ops.one_euro_filter(torch.as_tensor(p).cuda()).cpu().numpy()

    def smooth_points2d(self, cam_id, private_cache=dict()):
        """Smoothed pixel tracks of one camera (core.py:286-296 -> df3d/signal_util.py:135-160)."""
        import torch

        from . import ops

        if cam_id not in private_cache:
            pts = torch.as_tensor(np.ascontiguousarray(self.camNet.cam_list[cam_id].points2d, dtype=np.float64)).cuda()
            private_cache[cam_id] = ops.smooth_pose2d(pts).cpu().numpy()
        return private_cache[cam_id]

    def save(self):
        """Writes the result pickle (core.py:349-369)."""
        out = {"points2d": np.copy(self.points2d)}
        if self.camNet is not None and self.camNet.has_calibration():
            self.camNet.triangulate()
            pts3d = self.camNet.points3d
            out["points3d_wo_procrustes"] = pts3d
            out["points3d"] = self._procrustes_device(pts3d)
            out = {**self.camNet.summarize(), **out}
            self.points3d = out["points3d"]
        else:
            logger.debug("Triangulation skipped.")
        out["camera_ordering"] = self.camera_ordering
        out["heatmap_confidence"] = self.conf
        with open(self.save_path, "wb") as f:
            pickle.dump(out, f)
        print(f"Saved results at: {self.save_path}")

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _procrustes_device(pts3d):
        import torch

        from . import ops

        return ops.procrustes(torch.as_tensor(np.ascontiguousarray(pts3d, dtype=np.float64)).cuda()).cpu().numpy()

    def setup_camera_ordering(self, camera_ordering):
        if camera_ordering is None:
            camera_ordering = find_default_camera_ordering(self.input_folder)
        return np.array(camera_ordering)

    def expand_videos(self):
        """camera_x.mp4 -> camera_x_img_y.jpg when the images are missing (core.py:446-459).  With ffmpeg on the PATH
        this is the reference's command, byte for byte (`-qscale:v 2 -start_number 0`); without it OpenCV's bundled
        FFmpeg decodes and libjpeg writes at quality 95 -- same frames, slightly different JPEG quantisation."""
        import shutil
        import subprocess

        for vid in glob.glob(os.path.join(self.input_folder, "camera_?.mp4")):
            cam_id = int(re.match(r"camera_(\d+)", os.path.basename(vid))[1])
            if os.path.exists(os.path.join(self.input_folder, f"camera_{cam_id}_img_0.jpg")) or \
                    os.path.exists(os.path.join(self.input_folder, f"camera_{cam_id}_img_000000.jpg")):
                continue
            if shutil.which("ffmpeg"):
                subprocess.call(["ffmpeg", "-nostats", "-loglevel", "error", "-i", vid, "-qscale:v", "2", "-start_number", "0",
                                 os.path.join(self.input_folder, f"camera_{cam_id}_img_%d.jpg")], stdin=subprocess.DEVNULL)
                continue
            import cv2

            cap = cv2.VideoCapture(vid)
            i = 0
            ok, frame = cap.read()
            while ok:
                cv2.imwrite(os.path.join(self.input_folder, f"camera_{cam_id}_img_{i}.jpg"), frame,
                            [cv2.IMWRITE_JPEG_QUALITY, 95])
                i += 1
                ok, frame = cap.read()
            cap.release()

    def get_fps(self):
        """Frame rate of the input videos (core.py:416-444: ffprobe; OpenCV's container read when ffprobe is absent)."""
        import shutil
        import subprocess

        rates = []
        for vid in sorted(glob.glob(os.path.join(self.input_folder, "camera_?.mp4"))):
            if shutil.which("ffprobe"):
                try:
                    out = subprocess.check_output(["ffprobe", "-v", "error", "-select_streams", "v:0", "-show_entries",
                                                   "stream=avg_frame_rate", "-of", "default=noprint_wrappers=1:nokey=1", vid], text=True).strip()
                    num, _, den = out.partition("/")
                    rates.append(float(num) / float(den) if den and float(den) != 0 else float(num))
                    continue
                except Exception:
                    logger.warning(f"ffprobe failed on {vid}")
            import cv2

            cap = cv2.VideoCapture(vid)
            fps = cap.get(cv2.CAP_PROP_FPS)
            cap.release()
            if fps and fps > 0:
                rates.append(float(fps))
        if not rates:
            return None
        if any(abs(r - rates[0]) > 1e-9 for r in rates):
            logger.warning(f"Framerates of input videos differ from one another, using the first one: {rates}")
        return rates[0]

    def delete_images(self):
        """Deletes camera_N_img_*.jpg for every camera that has a camera_N.mp4 (core.py:461-475)."""
        for vid in glob.glob(os.path.join(self.input_folder, "camera_[0-9].mp4")):
            cam_id = int(re.match(r"camera_(\d+)", os.path.basename(vid))[1])
            for img in glob.glob(os.path.join(self.input_folder, f"camera_{cam_id}_img_*.jpg")):
                os.remove(img)
