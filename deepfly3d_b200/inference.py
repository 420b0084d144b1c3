"""df2d-compatible 2-D inference entry point backed by the CUDA hourglass.

``inference_folder`` keeps the signature and return convention of
``df2d.inference.inference_folder`` as called at df3d/core.py:177-185:

    points2d (7, T, 19, 2)  float64, (row / Hh, col / Wh) of the arg-max in the (possibly mirrored)
                            network frame
    conf     (7, T, 19, 1)  float32 peak value

Image ingest is a bounded-memory stream, like the reference's DataLoader: the
recording is cut into blocks of frames; a pool of host threads decodes block k+1 into pinned memory
(``cv2.imread`` releases the GIL) while a copy stream uploads block k and the main stream resizes it to the
network input (csrc/ingest.cu, bit-identical to ``cv2.resize(..., INTER_LINEAR)``) and runs the hourglass on it.
Host and device memory are bounded by two blocks, whatever the length of the recording.  By default the JPEG
files are decoded on the host (the reference's path too, bit-identical frames); ``gpu_decode=True`` decodes them
with nvJPEG on the device (a few grey levels away).
"""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .hourglass import HourglassEngine
from .skeleton import HEATMAP_SHAPE, NUM_CAMERAS, NUM_PREDICT

# images of one engine launch sequence at 256 x 256 (csrc/hourglass.cu: chunk_for); other input sizes scale by area
_CHUNK_IMAGES_256 = 1792


def image_name(folder, cam_id, img_id):
    plain = os.path.join(folder, f"camera_{cam_id}_img_{img_id}.jpg")
    if os.path.isfile(plain):
        return plain
    return os.path.join(folder, f"camera_{cam_id}_img_{img_id:06d}.jpg")


def plan_blocks(T, block_frames):
    """Frame blocks [(t0, t1), ...] covering [0, T) in order."""
    block_frames = max(1, int(block_frames))
    return [(t0, min(T, t0 + block_frames)) for t0 in range(0, T, block_frames)]


def block_frames_for(in_h, in_w, T, batch_size=8):
    """Frames per block: what one launch sequence of the engine takes (1 792 images at 256 x 256, scaled by the
    input area), never less than the reference's `batch_size` images, never more than the recording."""
    images = max(int(batch_size), (_CHUNK_IMAGES_256 * 256 * 256) // (int(in_h) * int(in_w)), NUM_CAMERAS)
    return max(1, min(int(T), images // NUM_CAMERAS))


def stream_block_frames(engine_block, gpu_decode=False):
    """Frames per streamed block when the caller did not choose: with the host decode -- the slower side of the stream --
    quarter-size blocks start the GPU earlier (the first block is the only one whose decode nothing hides) and still
    fill it (measured on 256 frames of 480x960 JPEGs, 16 host cores: 128-frame blocks 698, 64-frame 808, 32-frame 949
    frames/s through pose2d_estimation); the device decode takes whole engine blocks."""
    engine_block = int(engine_block)
    return engine_block // 4 if (not gpu_decode and engine_block >= 32) else engine_block


class FolderReader:
    """Threaded host-side read of ``camera_{c}_img_{t}.jpg`` frame blocks (native size, gray)."""

    def __init__(self, folder, workers=None):
        self.folder = folder
        self.workers = workers or min(32, os.cpu_count() or 1)
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="df3d-read")
        self._shape = None
        self._lock = threading.Lock()

    def close(self):
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def shape(self):
        """(H, W) of the recording's frames, from camera 0, image 0."""
        if self._shape is None:
            self._shape = tuple(self._decode(image_name(self.folder, 0, 0)).shape)
        return self._shape

    @staticmethod
    def _decode(path):
        import cv2

        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise FileNotFoundError(f"cannot read {path}")
        return img

    def _read_into(self, arr, c, t, slot):
        path = image_name(self.folder, c, t)
        img = self._decode(path)
        if img.shape != arr.shape[2:]:
            raise ValueError(f"{path}: image size {img.shape} differs from the first image {tuple(arr.shape[2:])}")
        arr[c, slot] = img

    def read_block_async(self, t0, t1, out):
        """Decodes frames [t0, t1) of all cameras into out[:, :t1-t0] ((7, >=t1-t0, H, W) uint8 tensor).
        Returns the futures; ``wait`` re-raises the first failure."""
        arr = out.numpy()
        return [self._pool.submit(self._read_into, arr, c, t, t - t0) for c in range(NUM_CAMERAS) for t in range(t0, t1)]

    def read_bytes_async(self, t0, t1):
        """Compressed streams of frames [t0, t1), camera-major (for the device-side decode)."""
        def rd(c, t):
            path = image_name(self.folder, c, t)
            if not os.path.isfile(path):
                raise FileNotFoundError(f"cannot read {path}")
            with open(path, "rb") as f:
                return f.read()
        return [self._pool.submit(rd, c, t) for c in range(NUM_CAMERAS) for t in range(t0, t1)]

    @staticmethod
    def wait(futures):
        return [f.result() for f in futures]


class VideoReader:
    """Frame blocks straight from ``camera_{c}.mp4`` (streaming ingest), without the
    reference's detour through JPEG files (``Core.expand_videos``, df3d/core.py:446-459).  One decode thread per
    camera (OpenCV's bundled FFmpeg; this image has no NVDEC binding -- no Video Codec SDK headers, PyNvVideoCodec,
    DALI or torchcodec), frames converted to gray on the host.  Blocks must be requested in order."""

    def __init__(self, folder, workers=None):
        import cv2

        self.folder = folder
        self.workers = NUM_CAMERAS
        self._caps = []
        for c in range(NUM_CAMERAS):
            path = os.path.join(folder, f"camera_{c}.mp4")
            cap = cv2.VideoCapture(path)
            if not cap.isOpened():
                raise FileNotFoundError(f"cannot open {path}")
            self._caps.append(cap)
        self.num_frames = min(int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) for cap in self._caps)
        self.shape = (int(self._caps[0].get(cv2.CAP_PROP_FRAME_HEIGHT)), int(self._caps[0].get(cv2.CAP_PROP_FRAME_WIDTH)))
        self._next = [0] * NUM_CAMERAS
        self._pool = ThreadPoolExecutor(max_workers=NUM_CAMERAS, thread_name_prefix="df3d-video")

    def close(self):
        self._pool.shutdown(wait=True)
        for cap in self._caps:
            cap.release()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _read_camera(self, arr, c, t0, t1):
        import cv2

        if self._next[c] != t0:
            raise ValueError(f"camera {c}: video blocks must be read in order (next frame {self._next[c]}, asked for {t0})")
        for t in range(t0, t1):
            ok, frame = self._caps[c].read()
            if not ok:
                raise FileNotFoundError(f"camera_{c}.mp4 ends at frame {t}")
            arr[c, t - t0] = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        self._next[c] = t1

    def read_block_async(self, t0, t1, out):
        arr = out.numpy()
        return [self._pool.submit(self._read_camera, arr, c, t0, t1) for c in range(NUM_CAMERAS)]

    @staticmethod
    def wait(futures):
        return [f.result() for f in futures]


def read_images(folder, max_img_id, pin_memory=True):
    """-> uint8 tensor (7, T, Hs, Ws): the gray frames at their native size in (pinned) host memory.  Whole
    recording at once -- small folders and tests; ``inference_folder`` streams blocks instead."""
    T = max_img_id + 1
    with FolderReader(folder) as rd:
        Hs, Ws = rd.shape
        out = torch.empty((NUM_CAMERAS, T, Hs, Ws), dtype=torch.uint8)
        if pin_memory and torch.cuda.is_available():
            out = out.pin_memory()
        rd.wait(rd.read_block_async(0, T, out))
    return out


def random_state_dict(num_stacks=2, num_classes=NUM_PREDICT, seed=0):
    """Seeded stand-in weights with the checkpoint's key layout (no pretrained weights offline)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, scale=1.0):
        sd[f"{name}.weight"] = torch.randn((co, ci, k, k), generator=g) * (scale / (ci * k * k) ** 0.5)
        sd[f"{name}.bias"] = torch.randn(co, generator=g) * 0.05

    def bn(name, n):
        sd[f"{name}.weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        sd[f"{name}.bias"] = 0.1 * torch.randn(n, generator=g)
        sd[f"{name}.running_mean"] = 0.1 * torch.randn(n, generator=g)
        sd[f"{name}.running_var"] = 1.0 + 0.1 * torch.rand(n, generator=g)

    def bott(name, inpl, planes):
        bn(f"{name}.bn1", inpl); conv(f"{name}.conv1", planes, inpl, 1)
        bn(f"{name}.bn2", planes); conv(f"{name}.conv2", planes, planes, 3)
        bn(f"{name}.bn3", planes); conv(f"{name}.conv3", 2 * planes, planes, 1, 0.3)
        if inpl != 2 * planes:
            conv(f"{name}.downsample.0", 2 * planes, inpl, 1)

    conv("conv1", 64, 3, 7); bn("bn1", 64)
    bott("layer1.0", 64, 64); bott("layer2.0", 128, 64); bott("layer3.0", 128, 128)
    for i in range(num_stacks):
        for d in range(4):
            for k in range(4 if d == 0 else 3):
                bott(f"hg.{i}.hg.{d}.{k}.0", 256, 128)
        bott(f"res.{i}.0", 256, 128)
        conv(f"fc.{i}.0", 256, 256, 1); bn(f"fc.{i}.1", 256)
        conv(f"score.{i}", num_classes, 256, 1)
        if i < num_stacks - 1:
            conv(f"fc_.{i}", 256, 256, 1, 0.1); conv(f"score_.{i}", 256, num_classes, 1, 0.1)
    return sd


def load_state_dict(weights=None):
    """Checkpoint path (torch file holding a state_dict, possibly under 'state_dict', keys possibly prefixed
    with DataParallel's 'module.' -- the layout of df2d's ``sh8_deepfly.tar``, reference df3d/config.py:30-32) or
    $DF3D_B200_WEIGHTS; without either, seeded stand-in weights are used and a warning is logged."""
    path = weights or os.environ.get("DF3D_B200_WEIGHTS")
    if path:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        return ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    import logging

    logging.getLogger("df3d.logger").warning(
        "no pretrained hourglass weights given (weights= / $DF3D_B200_WEIGHTS): using seeded random weights")
    return random_state_dict()


def load_mean(mean=None):
    """The per-channel mean subtracted from x/255: a number, three numbers, or the path of a torch file like
    df2d's ``mean.pth.tar`` (reference df3d/config.py:37-39: a dict with a 'mean' entry, or the bare tensor);
    None -> $DF3D_B200_MEAN if set, else 0.5."""
    if mean is None:
        mean = os.environ.get("DF3D_B200_MEAN")
        if mean is None:
            return (0.5, 0.5, 0.5)
    if isinstance(mean, (str, os.PathLike)):
        if os.path.isfile(mean):
            obj = torch.load(mean, map_location="cpu", weights_only=False)
            if isinstance(obj, dict):
                obj = obj["mean"]
            mean = obj
        else:
            mean = float(mean)
    m = np.asarray(mean.detach().cpu().numpy() if torch.is_tensor(mean) else mean, dtype=np.float64).reshape(-1)
    if m.size == 1:
        m = np.repeat(m, 3)
    if m.size != 3:
        raise ValueError(f"mean must have 1 or 3 entries, got {m.size}")
    return tuple(float(v) for v in m)


# ------------------------------------------------------------------------------------------------
# engine cache: ONE engine (its workspace is tens of GB), keyed on what the weights ARE, not on an object id
_ENGINE = {"key": None, "engine": None, "ref": None}
_ENGINE_LOCK = threading.Lock()


def _weights_key(state_dict, weights):
    if state_dict is not None:
        return ("dict", id(state_dict))      # the cache entry keeps the dict alive, so the id cannot be recycled
    path = weights or os.environ.get("DF3D_B200_WEIGHTS")
    if path:
        st = os.stat(path)
        return ("file", os.path.abspath(path), st.st_mtime_ns, st.st_size)
    return ("random", 0)


def get_engine(state_dict, in_h, in_w, max_batch, device="cuda", mean=(0.5, 0.5, 0.5), weights=None):
    """The cached engine for these weights / input size / mean, rebuilt when any of them changes or when a larger
    batch is asked for.  The previous engine is closed first (LRU of one)."""
    key = (_weights_key(state_dict, weights), int(in_h), int(in_w), str(device), tuple(mean))
    with _ENGINE_LOCK:
        eng = _ENGINE["engine"]
        if eng is not None and _ENGINE["key"] == key and eng.max_batch >= max_batch:
            return eng
        if eng is not None:
            eng.close()
            _ENGINE.update(key=None, engine=None, ref=None)
        sd = state_dict if state_dict is not None else load_state_dict(weights)
        eng = HourglassEngine(sd, in_h, in_w, max_batch, device=device, mean=mean)
        _ENGINE.update(key=key, engine=eng, ref=state_dict)
        return eng


def drop_engine():
    with _ENGINE_LOCK:
        if _ENGINE["engine"] is not None:
            _ENGINE["engine"].close()
        _ENGINE.update(key=None, engine=None, ref=None)


_JPEG = {}


def _jpeg_decoder(backend=None):
    from . import ops

    dec = _JPEG.get(backend)
    if dec is None:
        dec = _JPEG[backend] = ops.JpegDecoder(backend)
    return dec


def inference_folder(folder, camera_ids_to_flip=(), return_heatmap=False, return_confidence=True, max_img_id=None,
                     batch_size=8, disable_pin_memory=False, state_dict=None, weights=None, input_size=None,
                     device="cuda", gpu_decode=False, mean=None, block_frames=None, workers=None, stats=None, source="images"):
    """Runs the hourglass on every camera_{0..6}_img_{0..max_img_id}.jpg of `folder`.

    batch_size is the reference's DataLoader batch (images per forward).  Here a forward runs over a block of
    frames sized for the GPU (``block_frames_for``); a batch_size above that raises the block.  `stats`, when a
    dict, receives the block plan and the seconds spent waiting for the host decode.  source="videos" streams the
    frames from camera_{0..6}.mp4 instead of the expanded JPEG files (see VideoReader)."""
    from . import ops

    if max_img_id is None:
        raise ValueError("max_img_id is required")
    if not torch.cuda.is_available():
        raise RuntimeError("inference_folder needs a CUDA device (sm_100a); there is no CPU fallback")
    Hh, Wh = HEATMAP_SHAPE
    in_h, in_w = input_size if input_size is not None else (4 * Hh, 4 * Wh)
    T = max_img_id + 1
    bf = int(block_frames) if block_frames else stream_block_frames(block_frames_for(in_h, in_w, T, batch_size), gpu_decode)
    blocks = plan_blocks(T, bf)
    dev = torch.device(device)
    eng = get_engine(state_dict, in_h, in_w, NUM_CAMERAS * bf, device=device, mean=load_mean(mean), weights=weights)
    K = eng.num_classes
    hh, hw = eng.heatmap_shape
    flip_cam = torch.zeros(NUM_CAMERAS, dtype=torch.uint8)
    for c in camera_ids_to_flip:
        flip_cam[int(c)] = 1
    idx_all = torch.empty((NUM_CAMERAS, T, K), dtype=torch.int32, device=dev)
    conf_all = torch.empty((NUM_CAMERAS, T, K), dtype=torch.float32, device=dev)
    heat_all = torch.empty((NUM_CAMERAS, T, K, hh, hw), dtype=torch.float32) if return_heatmap else None
    import time

    wait_s = 0.0
    if source not in ("images", "videos") or (source == "videos" and gpu_decode):
        raise ValueError("source must be 'images' or 'videos' (videos are decoded on the host)")
    with torch.cuda.device(dev), (VideoReader(folder) if source == "videos" else FolderReader(folder, workers=workers)) as rd:
        main = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        if gpu_decode:
            # device-side decode: the files are read by the pool, the block is decoded on a side stream (hardware
            # JPEG engines where the GPU has them) while the hourglass of the previous block runs on the main one
            dec = _jpeg_decoder(gpu_decode if isinstance(gpu_decode, str) else None)
            dec_stream = torch.cuda.Stream()

            def submit_decode(futures):
                nonlocal wait_s
                t_w = time.perf_counter()
                data = rd.wait(futures)
                wait_s += time.perf_counter() - t_w
                with torch.cuda.stream(dec_stream):
                    frames = dec.decode_gray(data, device=dev)    # (7*tc, Hs, Ws), camera-major
                    ev = torch.cuda.Event()
                    ev.record(dec_stream)
                return frames, ev, data

            decoded = submit_decode(rd.read_bytes_async(*blocks[0]))
            pending = rd.read_bytes_async(*blocks[1]) if len(blocks) > 1 else None
        else:
            Hs, Ws = rd.shape
            host = [torch.empty((NUM_CAMERAS, bf, Hs, Ws), dtype=torch.uint8) for _ in range(2)]
            if not disable_pin_memory:
                host = [h.pin_memory() for h in host]
            staged = [torch.empty((NUM_CAMERAS, bf, Hs, Ws), dtype=torch.uint8, device=dev) for _ in range(2)]
            uploaded = [torch.cuda.Event() for _ in range(2)]   # H2D of the slot finished (host buffer reusable)
            consumed = [torch.cuda.Event() for _ in range(2)]   # compute done with the device slot
            for ev in consumed:
                ev.record(main)
            pending = rd.read_block_async(*blocks[0], host[0])
        for k, (t0, t1) in enumerate(blocks):
            tc = t1 - t0
            slot = k & 1
            if gpu_decode:
                native, ev, keep = decoded
                main.wait_event(ev)
                native.record_stream(main)
            else:
                t_w = time.perf_counter()
                rd.wait(pending)                                 # block k decoded
                wait_s += time.perf_counter() - t_w
                if k + 1 < len(blocks):
                    if k >= 1:
                        uploaded[slot ^ 1].synchronize()         # the other host buffer has left for the device
                    pending = rd.read_block_async(*blocks[k + 1], host[slot ^ 1])
                copy_stream.wait_event(consumed[slot])
                with torch.cuda.stream(copy_stream):
                    staged[slot][:, :tc].copy_(host[slot][:, :tc], non_blocking=True)
                    uploaded[slot].record(copy_stream)
                main.wait_event(uploaded[slot])
                native = staged[slot][:, :tc].reshape(NUM_CAMERAS * tc, Hs, Ws) if tc == bf else \
                    staged[slot][:, :tc].contiguous().reshape(NUM_CAMERAS * tc, Hs, Ws)
            images = native if tuple(native.shape[1:]) == (in_h, in_w) else ops.resize_gray_u8(native, (in_h, in_w))
            flip = flip_cam.view(NUM_CAMERAS, 1).expand(NUM_CAMERAS, tc).reshape(-1).to(dev)
            res = eng.forward(images, flip=flip, return_heatmap=return_heatmap)
            idx_all[:, t0:t1] = res[0].view(NUM_CAMERAS, tc, K)
            conf_all[:, t0:t1] = res[1].view(NUM_CAMERAS, tc, K)
            if not gpu_decode:
                consumed[slot].record(main)
            else:
                if pending is not None:                          # decode of block k+1 overlaps the forward of block k
                    decoded = submit_decode(pending)
                    pending = rd.read_bytes_async(*blocks[k + 2]) if k + 2 < len(blocks) else None
                ev.synchronize()                                 # block k is decoded (it ran ahead of its forward):
                keep = None                                      # its compressed streams may go
            if return_heatmap:
                heat_all[:, t0:t1] = res[2][..., :K].permute(0, 3, 1, 2).reshape(NUM_CAMERAS, tc, K, hh, hw).cpu()
        idx_h = idx_all.cpu().numpy().astype(np.int64)
        conf_h = conf_all.cpu().numpy()
    if stats is not None:
        stats.update(blocks=len(blocks), block_frames=bf, decode_wait_s=wait_s, workers=rd.workers,
                     decode=("nvjpeg-" + dec.backend) if gpu_decode else "host")
    points2d = np.stack([(idx_h // hw) / hh, (idx_h % hw) / hw], axis=-1).astype(np.float64)
    out = [points2d]
    if return_heatmap:
        out.append(heat_all.numpy())
    if return_confidence:
        out.append(conf_h.reshape(NUM_CAMERAS, T, -1, 1))
    return tuple(out) if len(out) > 1 else out[0]
