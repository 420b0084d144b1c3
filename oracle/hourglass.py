"""Oracle for the df2d half of the path: stacked-hourglass forward (PyTorch, CPU).

Test infrastructure (see ``oracle/__init__.py``).  **Parity unpinned**: df2d is not vendored in
``/root/reference`` and the pretrained ``sh8_deepfly.tar`` weights (reference
``df3d/config.py:30-32``) are unavailable, so the golden 2-D points of
``tests/test_df3d.py:150-196`` cannot be regenerated here.  The architecture is the published
stacked hourglass (Newell et al., ECCV 2016; pre-activation bottlenecks, expansion 2,
num_feats 128, depth-4 hourglass, nearest x2 up-sampling) with the hyper-parameters hinted by
``df3d/config.py:18,33-36`` (19 output maps, heat-map = input / 4).  Module names follow the
public ``pytorch-pose`` hourglass so a real checkpoint would load by name.  Its FLOP count
reproduces SURVEY.md section 8(d) (54.975 GFLOP / 256x256 image for 8 stacks) -- see
``tests/test_oracle_hourglass.py``.

``emulate_bf16=True`` rounds exactly where the CUDA path rounds (every tensor that is written
to HBM is bf16; conv weights are bf16; accumulation, bias and BN arithmetic are fp32), so the
arg-max indices of kernel and oracle can be compared on the same arithmetic.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Bottleneck(nn.Module):
    expansion = 2

    def __init__(self, inplanes, planes):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=True)
        self.bn3 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 2, 1, bias=True)
        self.downsample = None
        if inplanes != planes * 2:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 2, 1, bias=True))

    def forward(self, x, rnd=lambda t: t, wq=lambda c: c.weight, round_shortcut=True):
        """round_shortcut=False (emulation only): the CUDA path evaluates conv3(t2) + downsample(x) of layer1 as one
        GEMM over the concatenated inputs, so the projection shortcut is never rounded to bf16 on its own."""
        a = rnd(F.relu(self.bn1(x)))
        t1 = rnd(F.relu(self.bn2(F.conv2d(a, wq(self.conv1), self.conv1.bias))))
        t2 = rnd(F.relu(self.bn3(F.conv2d(t1, wq(self.conv2), self.conv2.bias, padding=1))))
        res = x
        if self.downsample is not None:
            ds = self.downsample[0]
            res = F.conv2d(x, wq(ds), ds.bias)
            if round_shortcut:
                res = rnd(res)
        return rnd(F.conv2d(t2, wq(self.conv3), self.conv3.bias) + res)


class Hourglass(nn.Module):
    def __init__(self, planes, depth):
        super().__init__()
        self.depth = depth
        hg = []
        for i in range(depth):
            res = [nn.Sequential(Bottleneck(planes * 2, planes)) for _ in range(3)]
            if i == 0:
                res.append(nn.Sequential(Bottleneck(planes * 2, planes)))
            hg.append(nn.ModuleList(res))
        self.hg = nn.ModuleList(hg)

    def _fwd(self, n, x, rnd, wq):
        up1 = self.hg[n - 1][0][0](x, rnd, wq)
        low1 = F.max_pool2d(x, 2, stride=2)
        low1 = self.hg[n - 1][1][0](low1, rnd, wq)
        if n > 1:
            low2 = self._fwd(n - 1, low1, rnd, wq)
        else:
            low2 = self.hg[n - 1][3][0](low1, rnd, wq)
        low3 = self.hg[n - 1][2][0](low2, rnd, wq)
        up2 = F.interpolate(low3, scale_factor=2, mode="nearest")
        return rnd(up1 + up2)

    def forward(self, x, rnd=lambda t: t, wq=lambda c: c.weight):
        return self._fwd(self.depth, x, rnd, wq)


class HourglassNet(nn.Module):
    """Stacked hourglass: (B,3,H,W) -> list of num_stacks score maps (B,K,H/4,W/4)."""

    def __init__(self, num_stacks=2, num_classes=19, num_feats=128, inplanes=64):
        super().__init__()
        self.num_stacks, self.num_classes = num_stacks, num_classes
        ch = num_feats * 2
        self.conv1 = nn.Conv2d(3, inplanes, 7, stride=2, padding=3, bias=True)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.layer1 = nn.Sequential(Bottleneck(inplanes, inplanes))            # 64 -> 128 @ H/2
        self.layer2 = nn.Sequential(Bottleneck(inplanes * 2, inplanes))        # 128 -> 128 @ H/4
        self.layer3 = nn.Sequential(Bottleneck(inplanes * 2, num_feats))       # 128 -> 256 @ H/4
        self.hg = nn.ModuleList([Hourglass(num_feats, 4) for _ in range(num_stacks)])
        self.res = nn.ModuleList([nn.Sequential(Bottleneck(ch, num_feats)) for _ in range(num_stacks)])
        self.fc = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(ch, ch, 1, bias=True), nn.BatchNorm2d(ch), nn.ReLU(inplace=True))
             for _ in range(num_stacks)])
        self.score = nn.ModuleList([nn.Conv2d(ch, num_classes, 1, bias=True) for _ in range(num_stacks)])
        self.fc_ = nn.ModuleList([nn.Conv2d(ch, ch, 1, bias=True) for _ in range(num_stacks - 1)])
        self.score_ = nn.ModuleList([nn.Conv2d(num_classes, ch, 1, bias=True) for _ in range(num_stacks - 1)])

    def forward(self, x, emulate_bf16=False, gray_fold=None, fc_merge=True):
        """gray_fold (emulation only): the CUDA path takes uint8 gray images whose three input planes are
        identical and folds the stem weights over the input channel (fp32 sum, one bf16 rounding);
        None = do the same whenever the three planes of `x` are identical.
        fc_merge (emulation only): the CUDA path evaluates fc(res.conv3(t) + h) as (W_fc W_3) t + W_fc h -- the
        product merged in fp64 and rounded to bf16 once, the sum conv3(t) + h never rounded (DF3D_HG_NO_FC_MERGE=1
        restores the separate convolutions, fc_merge=False emulates that)."""
        rnd = _bf16 if emulate_bf16 else (lambda t: t)
        wq = (lambda c: _bf16(c.weight)) if emulate_bf16 else (lambda c: c.weight)
        x = rnd(x)
        if gray_fold is None:
            gray_fold = bool(torch.equal(x[:, 0], x[:, 1]) and torch.equal(x[:, 0], x[:, 2]))
        if emulate_bf16 and gray_fold:
            w1 = _bf16(self.conv1.weight.sum(dim=1, keepdim=True))
            x = rnd(F.relu(self.bn1(F.conv2d(x[:, :1], w1, self.conv1.bias, stride=2, padding=3))))
        else:
            x = rnd(F.relu(self.bn1(F.conv2d(x, wq(self.conv1), self.conv1.bias, stride=2, padding=3))))
        x = self.layer1[0](x, rnd, wq, round_shortcut=not emulate_bf16)
        x = F.max_pool2d(x, 2, stride=2)
        x = self.layer2[0](x, rnd, wq)
        x = self.layer3[0](x, rnd, wq)
        out = []
        for i in range(self.num_stacks):
            y = self.hg[i](x, rnd, wq)
            fc_conv, fc_bn = self.fc[i][0], self.fc[i][1]
            if emulate_bf16 and fc_merge:
                rb = self.res[i][0]
                a = rnd(F.relu(rb.bn1(y)))
                t1 = rnd(F.relu(rb.bn2(F.conv2d(a, wq(rb.conv1), rb.conv1.bias))))
                t2 = rnd(F.relu(rb.bn3(F.conv2d(t1, wq(rb.conv2), rb.conv2.bias, padding=1))))
                Wfc, W3 = fc_conv.weight.double()[:, :, 0, 0], rb.conv3.weight.double()[:, :, 0, 0]
                Wm = (Wfc @ W3).float()
                bm = (Wfc @ rb.conv3.bias.double() + fc_conv.bias.double()).float()
                pre = F.conv2d(t2, _bf16(Wm)[:, :, None, None]) + F.conv2d(y, wq(fc_conv)) + bm.view(1, -1, 1, 1)
                y = rnd(F.relu(fc_bn(pre)))
            else:
                y = self.res[i][0](y, rnd, wq)
                y = rnd(F.relu(fc_bn(F.conv2d(y, wq(fc_conv), fc_conv.bias))))
            score = F.conv2d(y, wq(self.score[i]), self.score[i].bias)  # fp32, not rounded
            out.append(score)
            if i < self.num_stacks - 1:
                if emulate_bf16:
                    # the CUDA path merges the three linear maps x + fc_(y) + score_(score(y)) into
                    # one 1x1 conv (weights merged in fp32/fp64, then rounded to bf16 once)
                    Ws, bs = self.score[i].weight.double()[:, :, 0, 0], self.score[i].bias.double()
                    Wr, br = self.score_[i].weight.double()[:, :, 0, 0], self.score_[i].bias.double()
                    Wm = (self.fc_[i].weight.double()[:, :, 0, 0] + Wr @ Ws).float()
                    bm = (self.fc_[i].bias.double() + Wr @ bs + br).float()
                    x = rnd(F.conv2d(y, _bf16(Wm)[:, :, None, None], bm) + x)
                else:
                    x = x + F.conv2d(y, self.fc_[i].weight, self.fc_[i].bias) \
                        + F.conv2d(score, self.score_[i].weight, self.score_[i].bias)
        return out


def make_model(num_stacks=2, num_classes=19, seed=0, res_scale=0.3, skip_scale=0.1):
    """Seeded random weights (no pretrained weights are available): 1/sqrt(fan_in) normal convs,
    residual branches (conv3) scaled by `res_scale` and the inter-stack re-injection convs
    (fc_, score_) by `skip_scale` so activations stay bounded through 8 stacks; BN running stats
    perturbed around (0, 1) so that BN folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    m = HourglassNet(num_stacks=num_stacks, num_classes=num_classes)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * (1.0 / fan_in) ** 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
            elif isinstance(mod, nn.BatchNorm2d):
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(1.0 + 0.1 * torch.rand(mod.running_var.shape, generator=g))
        for mod in m.modules():
            if isinstance(mod, Bottleneck):
                mod.conv3.weight.mul_(res_scale)
        for c in list(m.fc_) + list(m.score_):
            c.weight.mul_(skip_scale)
    m.eval()
    return m


def conv_flops(model, in_h, in_w):
    """2*MAC over all conv layers for one image (BN/ReLU/pool/upsample/add count 0)."""
    total = [0]
    hooks = []

    def hook(mod, inp, out):
        k = mod.kernel_size[0] * mod.kernel_size[1]
        total[0] += 2 * mod.in_channels * mod.out_channels * k * out.shape[2] * out.shape[3]

    # F.conv2d is called functionally above, so count through a shadow forward with module calls
    def counting_conv(x, w, b=None, stride=1, padding=0):
        out = _orig(x, w, b, stride=stride, padding=padding)
        total[0] += 2 * w.shape[1] * w.shape[0] * w.shape[2] * w.shape[3] * out.shape[2] * out.shape[3]
        return out

    _orig = F.conv2d
    F.conv2d = counting_conv
    try:
        with torch.no_grad():
            model(torch.zeros(1, 3, in_h, in_w))
    finally:
        F.conv2d = _orig
        for h in hooks:
            h.remove()
    return total[0]


@torch.no_grad()
def synthetic_images(n, h, w, seed=0, n_blobs=19, sigma=6.0):
    """SURVEY 8(d) config 2 input: sum of Gaussian blobs + N(0, 0.05) noise, gray, in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    cy = torch.rand(n, n_blobs, 1, 1, generator=g) * h
    cx = torch.rand(n, n_blobs, 1, 1, generator=g) * w
    img = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sigma * sigma)).sum(1)
    img = img + 0.05 * torch.randn(n, h, w, generator=g)
    return img.clamp_(0, 1)


def to_uint8(img01):
    return (img01 * 255.0).round().to(torch.uint8)


def preprocess_u8(img_u8, flip=None, mean=0.5):
    """uint8 gray (B,H,W) -> (B,3,H,W) float: x/255 - mean, replicated; optional LR mirror per image."""
    x = img_u8.to(torch.float32) / 255.0 - mean
    if flip is not None:
        fl = torch.as_tensor(flip, dtype=torch.bool)
        x = torch.where(fl.view(-1, 1, 1), x.flip(-1), x)
    return x.unsqueeze(1).expand(-1, 3, -1, -1).contiguous()
