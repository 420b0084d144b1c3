"""df2d-compatible 2-D inference entry point backed by the CUDA hourglass.

``inference_folder`` keeps the signature and return convention of
``df2d.inference.inference_folder`` as called at df3d/core.py:177-185:

    points2d (7, T, 19, 2)  float64, (row / Hh, col / Wh) of the arg-max in the (possibly mirrored)
                            network frame
    conf     (7, T, 19, 1)  float32 peak value

Image ingest: by default the JPEG files are decoded on the host (the
reference's path too, bit-identical frames); `gpu_decode=True` decodes them with nvJPEG on the device (a few
grey levels away).  The resize to the network input always runs on the device (csrc/ingest.cu, bit-identical
to the cv2.resize(..., INTER_LINEAR) the loader used to do on the host).
"""
import os

import numpy as np
import torch

from .hourglass import HourglassEngine
from .skeleton import HEATMAP_SHAPE, NUM_CAMERAS, NUM_PREDICT

_ENGINES = {}


def image_name(folder, cam_id, img_id):
    plain = os.path.join(folder, f"camera_{cam_id}_img_{img_id}.jpg")
    if os.path.isfile(plain):
        return plain
    return os.path.join(folder, f"camera_{cam_id}_img_{img_id:06d}.jpg")


def read_images(folder, max_img_id, pin_memory=True):
    """-> uint8 tensor (7, T, Hs, Ws): the gray frames at their native size, in pinned host memory."""
    import cv2

    T = max_img_id + 1
    out = None
    for c in range(NUM_CAMERAS):
        for t in range(T):
            path = image_name(folder, c, t)
            img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
            if img is None:
                raise FileNotFoundError(f"cannot read {path}")
            if out is None:
                out = torch.empty((NUM_CAMERAS, T) + img.shape, dtype=torch.uint8)
                if pin_memory and torch.cuda.is_available():
                    out = out.pin_memory()
                arr = out.numpy()
            if img.shape != tuple(out.shape[2:]):
                raise ValueError(f"{path}: image size {img.shape} differs from the first image {tuple(out.shape[2:])}")
            arr[c, t] = img
    return out


_JPEG = {}


def decode_images_device(folder, max_img_id, device="cuda"):
    """-> uint8 tensor (7*T, Hs, Ws) ON THE DEVICE: the compressed files are read on the host and decoded by
    nvJPEG (ops.JpegDecoder).  Opt-in: a few grey levels away from the host's libjpeg read."""
    from . import ops

    dec = _JPEG.get("dec")
    if dec is None:
        dec = _JPEG["dec"] = ops.JpegDecoder()
    streams = []
    for c in range(NUM_CAMERAS):
        for t in range(max_img_id + 1):
            path = image_name(folder, c, t)
            if not os.path.isfile(path):
                raise FileNotFoundError(f"cannot read {path}")
            with open(path, "rb") as f:
                streams.append(f.read())
    return dec.decode_gray(streams, device=device)


def load_images(folder, max_img_id, size_hw, pin_memory=True, device="cuda", gpu_decode=False):
    """-> uint8 tensor (7*T, H, W) gray ON THE DEVICE, resized to the network input (camera-major)."""
    from . import ops

    if gpu_decode:
        dev = decode_images_device(folder, max_img_id, device=device)
        Hs, Ws = dev.shape[1:]
    else:
        native = read_images(folder, max_img_id, pin_memory=pin_memory)
        C, T, Hs, Ws = native.shape
        dev = native.reshape(C * T, Hs, Ws).to(device, non_blocking=True)
    if (Hs, Ws) != tuple(size_hw):
        dev = ops.resize_gray_u8(dev, size_hw)
    return dev


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
    """Checkpoint path (torch file holding a state_dict, possibly under 'state_dict') or
    $DF3D_B200_WEIGHTS; without either, seeded stand-in weights are used and a warning is logged."""
    path = weights or os.environ.get("DF3D_B200_WEIGHTS")
    if path:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        return ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
    import logging

    logging.getLogger("df3d.logger").warning(
        "no pretrained hourglass weights given (weights= / $DF3D_B200_WEIGHTS): using seeded random weights")
    return random_state_dict()


def get_engine(state_dict, in_h, in_w, max_batch, device="cuda"):
    key = (id(state_dict), in_h, in_w, device)
    eng = _ENGINES.get(key)
    if eng is None or eng.max_batch < max_batch:
        eng = HourglassEngine(state_dict, in_h, in_w, max_batch, device=device)
        _ENGINES[key] = eng
    return eng


def inference_folder(folder, camera_ids_to_flip=(), return_heatmap=False, return_confidence=True, max_img_id=None,
                     batch_size=8, disable_pin_memory=False, state_dict=None, weights=None, input_size=None,
                     device="cuda", gpu_decode=False):
    """Runs the hourglass on every camera_{0..6}_img_{0..max_img_id}.jpg of `folder`."""
    if max_img_id is None:
        raise ValueError("max_img_id is required")
    Hh, Wh = HEATMAP_SHAPE
    in_h, in_w = input_size if input_size is not None else (4 * Hh, 4 * Wh)
    T = max_img_id + 1
    dev_images = load_images(folder, max_img_id, (in_h, in_w), pin_memory=not disable_pin_memory, device=device,
                             gpu_decode=gpu_decode)
    sd = state_dict if state_dict is not None else load_state_dict(weights)
    # batch_size is the reference's DataLoader batch; here the whole folder is one device batch and
    # the engine chunks internally, so it only bounds the workspace for tiny folders
    eng = get_engine(sd, in_h, in_w, max(NUM_CAMERAS * T, batch_size), device=device)
    flip = torch.zeros((NUM_CAMERAS, T), dtype=torch.uint8)
    for c in camera_ids_to_flip:
        flip[int(c)] = 1
    res = eng.forward(dev_images, flip=flip.reshape(-1).to(device), return_heatmap=return_heatmap)
    idx, conf = res[0], res[1]
    idx_h = idx.cpu().numpy().astype(np.int64).reshape(NUM_CAMERAS, T, -1)
    hh, hw = eng.heatmap_shape
    points2d = np.stack([(idx_h // hw) / hh, (idx_h % hw) / hw], axis=-1).astype(np.float64)
    out = [points2d]
    if return_heatmap:
        K = eng.num_classes
        out.append(res[2][..., :K].permute(0, 3, 1, 2).reshape(NUM_CAMERAS, T, K, hh, hw).cpu().numpy())
    if return_confidence:
        out.append(conf.cpu().numpy().reshape(NUM_CAMERAS, T, -1, 1))
    return tuple(out) if len(out) > 1 else out[0]
