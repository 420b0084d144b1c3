"""Host side of the stacked-hourglass engine: parameter flattening + handle management.

The CUDA library takes ONE flat float32 blob whose order is fixed by ``flatten_state_dict`` (it
mirrors ``read_net`` in csrc/hourglass.cu).  State-dict keys follow the public pytorch-pose
hourglass naming, which is what a df2d checkpoint (``sh8_deepfly.tar``, reference
``df3d/config.py:30-32``) uses: conv1, bn1, layer{1,2,3}.0.*, hg.S.hg.D.K.0.*, res.S.0.*,
fc.S.{0,1}.*, score.S.*, fc_.S.*, score_.S.*.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

HEAT_PAD = 32  # fp32 score channels stored per pixel by the kernels


def _bn(sd, prefix):
    return [sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"]]


def _conv(sd, prefix):
    return [sd[f"{prefix}.weight"], sd[f"{prefix}.bias"]]


def _bottleneck(sd, prefix):
    out = []
    out += _bn(sd, f"{prefix}.bn1") + _conv(sd, f"{prefix}.conv1")
    out += _bn(sd, f"{prefix}.bn2") + _conv(sd, f"{prefix}.conv2")
    out += _bn(sd, f"{prefix}.bn3") + _conv(sd, f"{prefix}.conv3")
    if f"{prefix}.downsample.0.weight" in sd:
        out += _conv(sd, f"{prefix}.downsample.0")
    return out


def infer_num_stacks(sd):
    n = 0
    while f"score.{n}.weight" in sd:
        n += 1
    if n == 0:
        raise ValueError("state dict has no score.N.weight entries: not an hourglass checkpoint")
    return n


def flatten_state_dict(sd, num_stacks=None):
    """state_dict -> (float32 numpy blob, num_stacks, num_classes)."""
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}  # DataParallel prefix
    S = infer_num_stacks(sd) if num_stacks is None else num_stacks
    K = sd["score.0.weight"].shape[0]
    t = []
    t += _conv(sd, "conv1") + _bn(sd, "bn1")
    for name in ("layer1.0", "layer2.0", "layer3.0"):
        t += _bottleneck(sd, name)
    for i in range(S):
        for d in range(4):
            for k in range(4 if d == 0 else 3):
                t += _bottleneck(sd, f"hg.{i}.hg.{d}.{k}.0")
        t += _bottleneck(sd, f"res.{i}.0")
        t += _conv(sd, f"fc.{i}.0") + _bn(sd, f"fc.{i}.1")
        t += _conv(sd, f"score.{i}")
        if i < S - 1:
            t += _conv(sd, f"fc_.{i}") + _conv(sd, f"score_.{i}")
    blob = np.concatenate([np.asarray(x.detach().cpu().float().numpy()).ravel() for x in t]).astype(np.float32)
    return blob, S, K


class HourglassEngine:
    """Owns a df3d_hg handle and its workspace.  ``forward`` takes CUDA tensors:

    * uint8 gray ``(B,H,W)`` (normalised on the fly as x/255 - mean and replicated to 3 channels), or
    * float32 ``(B,3,H,W)`` already normalised,

    and returns ``idx (B,K) int32`` (flat arg-max index in the H/4 x W/4 heat-map) and
    ``conf (B,K) float32`` of the LAST stack, optionally the fp32 heat-maps ``(B,H/4,W/4,32)``.
    """

    def __init__(self, state_dict, in_h, in_w, max_batch, device="cuda", mean=0.5):
        if not torch.cuda.is_available():
            raise RuntimeError("HourglassEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        blob, S, K = flatten_state_dict(state_dict)
        self.num_stacks, self.num_classes = S, K
        self.in_h, self.in_w, self.max_batch = in_h, in_w, max_batch
        self.device = torch.device(device)
        self.desc = _lib.HGDesc(S, K, in_h, in_w, max_batch)
        need = lib.df3d_hg_param_count(C.byref(self.desc))
        if need != blob.size:
            raise ValueError(f"state dict has {blob.size} parameters, the architecture needs {need}")
        self._h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            check(lib.df3d_hg_create(C.byref(self.desc), blob.ctypes.data_as(C.c_void_p), blob.size, C.byref(self._h)))
            m = (mean, mean, mean) if np.isscalar(mean) else tuple(mean)
            check(lib.df3d_hg_set_mean(self._h, *[float(v) for v in m]))
            self.ws_bytes = lib.df3d_hg_workspace_bytes(C.byref(self.desc))
            self.workspace = torch.empty(self.ws_bytes + 1024, dtype=torch.uint8, device=self.device)

    @property
    def heatmap_shape(self):
        return (self.in_h // 4, self.in_w // 4)

    def launches(self, B):
        return lib.df3d_hg_launches_per_forward(self._h, B)

    def forward(self, images, flip=None, return_heatmap=False):
        if not images.is_cuda:
            raise ValueError("HourglassEngine.forward takes CUDA tensors")
        images = images.contiguous()
        if images.dtype == torch.uint8 and images.dim() == 3:
            dtype, B = 0, images.shape[0]
            hw = tuple(images.shape[1:])
        elif images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3:
            dtype, B = 1, images.shape[0]
            hw = tuple(images.shape[2:])
        else:
            raise ValueError("images must be uint8 (B,H,W) or float32 (B,3,H,W)")
        if hw != (self.in_h, self.in_w):
            raise ValueError(f"engine was built for {self.in_h}x{self.in_w} images, got {hw}")
        if flip is not None:
            flip = flip.to(device=images.device, dtype=torch.uint8).contiguous()
            if flip.numel() != B:
                raise ValueError("flip must have one entry per image")
        K = self.num_classes
        idx = torch.empty((B, K), dtype=torch.int32, device=images.device)
        conf = torch.empty((B, K), dtype=torch.float32, device=images.device)
        Hh, Wh = self.heatmap_shape
        heat = torch.empty((B, Hh, Wh, HEAT_PAD), dtype=torch.float32, device=images.device) if return_heatmap else None
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.df3d_hg_forward_argmax(
            self._h, C.c_void_p(images.data_ptr()), dtype, C.c_void_p(flip.data_ptr() if flip is not None else 0), B,
            C.c_void_p(idx.data_ptr()), C.c_void_p(conf.data_ptr()), C.c_void_p(heat.data_ptr() if heat is not None else 0),
            C.c_void_p(self.workspace.data_ptr()), self.workspace.numel(), stream))
        if return_heatmap:
            return idx, conf, heat
        return idx, conf

    def set_timing(self, enable):
        check(lib.df3d_hg_set_timing(self._h, int(bool(enable))))

    def read_timing(self):
        """Per-kernel-class device times of the last forward (see df3d_hg_read_timing)."""
        out = (C.c_double * 8)()
        check(lib.df3d_hg_read_timing(self._h, out))
        keys = ("conv_ms", "conv_flop", "conv_launches", "other_ms", "other_launches", "conv3x3_ms", "conv3x3_flop",
                "conv3x3_launches")
        return dict(zip(keys, list(out)))

    def op_table(self):
        """[(label, ms, flop, bytes)] per plan entry for the last timed forward."""
        rows = []
        buf = C.create_string_buffer(128)
        for i in range(lib.df3d_hg_num_ops(self._h)):
            ms, fl, by = C.c_double(), C.c_double(), C.c_double()
            check(lib.df3d_hg_op_timing(self._h, i, C.byref(ms), C.byref(fl), C.byref(by), buf, 128))
            rows.append((buf.value.decode(), ms.value, fl.value, by.value))
        return rows

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.df3d_hg_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def conv2d_nhwc_bf16(x, weight, scale1, shift1, relu1=False, residual=None, scale2=None, shift2=None):
    """Operator-level entry (tests): x (B,H,W,Cin) bf16 CUDA, weight (Cout,Cin,k,k) float32 CPU.
    Returns (out_raw, out_act-or-None), both (B,H,W,Cout) bf16."""
    B, H, W, Cin = x.shape
    Cout, _, k, _ = weight.shape
    w = np.ascontiguousarray(weight.detach().cpu().float().numpy())
    s1 = np.ascontiguousarray(scale1.detach().cpu().float().numpy())
    h1 = np.ascontiguousarray(shift1.detach().cpu().float().numpy())
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=x.device)
    act = None
    s2p = h2p = C.c_void_p(0)
    if scale2 is not None:
        s2 = np.ascontiguousarray(scale2.detach().cpu().float().numpy())
        h2 = np.ascontiguousarray(shift2.detach().cpu().float().numpy())
        s2p, h2p = s2.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p)
        act = torch.empty_like(out)
    x = x.contiguous()
    if residual is not None:
        residual = residual.contiguous()
    check(lib.df3d_conv2d_nhwc_bf16(
        C.c_void_p(x.data_ptr()), B, H, W, Cin, w.ctypes.data_as(C.c_void_p), Cout, k,
        s1.ctypes.data_as(C.c_void_p), h1.ctypes.data_as(C.c_void_p), int(relu1),
        C.c_void_p(residual.data_ptr() if residual is not None else 0), C.c_void_p(out.data_ptr()), s2p, h2p,
        C.c_void_p(act.data_ptr() if act is not None else 0), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out, act
