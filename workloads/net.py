"""BASELINE.json configs[4] ("end-to-end"): the activations a ZebraPose network hands to the pose path.

The network body is NOT part of the path: it is stock library convolutions (cuDNN through torch).
What the path needs from it are the two tensors the reference concatenates in front of its last layer -- `x` (256
channels) and `x_128` (64 channels), both at 128 x 128 (model/aspp.py:112) -- plus that layer's weights for
`Engine.upload_head`.  This module is the feeder: a plain torch definition of the `BinaryCodeNet_Deeplab(34, 16, 2,
concat=True, output_kernel_size=1)` graph (model/BinaryCodeNet.py:122-177, model/resnet.py:160-247, model/aspp.py:5-114)
that stops in front of `conv_1x1_4`, in bf16 / channels_last with the BatchNorms folded into the convolutions.

Parameter names follow the reference's state dict (a trained ZebraPose checkpoint is the wire format here), so
`load_reference_state_dict` takes a reference checkpoint as it is and `reference_state_dict` writes one the reference
model loads with strict=True -- that is how tests/golden/make_golden_net.py pins this graph against the reference's own
forward pass.  No pretrained weights exist offline: `build(seed)` is the random-init network of configs[4].
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

STAGES = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2), (512, 3, 1, 4))     # (channels, blocks, stride, dilation)


def _cb(conv, bn, x, relu=True):
    y = bn(conv(x))
    return F.relu(y, inplace=True) if relu else y


def tap_conv3x3(conv, x):
    """A dilated 3x3 convolution as its nine taps: tap (ky, kx) is a 1x1 convolution of the input shifted by
    ((ky-1) d, (kx-1) d), and it only touches the output pixels whose shifted source lies inside the map.  On the 32 x 32
    ASPP map with d = 18 the off-centre taps cover 14 of 32 rows / columns, so 3.5 of the 9 taps' products remain -- and
    cuDNN's channels_last bf16 kernel for that shape is pathological (225 ms at 128 crops vs 0.25 ms for d = 6 / 12,
    tools/prof_net_body.py).  Same sums as F.conv2d in a different order."""
    d, (H, W) = conv.dilation[0], x.shape[2:]
    assert conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (d, d) and conv.dilation == (d, d)
    w = conv.weight
    out = F.conv2d(x, w[:, :, 1:2, 1:2], conv.bias)
    for ky in range(3):
        for kx in range(3):
            dy, dx = (ky - 1) * d, (kx - 1) * d
            y0, y1, x0, x1 = max(0, -dy), min(H, H - dy), max(0, -dx), min(W, W - dx)
            if (ky, kx) == (1, 1) or y0 >= y1 or x0 >= x1:
                continue
            out[:, :, y0:y1, x0:x1] += F.conv2d(x[:, :, y0 + dy:y1 + dy, x0 + dx:x1 + dx], w[:, :, ky:ky + 1, kx:kx + 1])
    return out


class Residual(nn.Module):
    """3x3 -> 3x3 residual unit; keys conv1/bn1/conv2/bn2/downsample.{0,1} as torchvision's and model/resnet.py:20-51."""

    def __init__(self, cin, cout, stride, dilation):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, dilation, dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, dilation, dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = nn.Sequential()
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = _cb(self.conv2, self.bn2, _cb(self.conv1, self.bn1, x), relu=False)
        y += self.downsample(x)
        return F.relu(y, inplace=True)


def _stage(cin, spec):
    c, n, s, d = spec
    return nn.Sequential(*[Residual(cin if i == 0 else c, c, s if i == 0 else 1, d) for i in range(n)])


class Backbone(nn.Module):
    """ResNet34 at output stride 8 with the three skip taps of the concat decoder (model/resnet.py:177-247)."""

    def __init__(self):
        super().__init__()
        self.resnet_layer_1 = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        self.resnet_layer_2 = nn.Sequential(nn.MaxPool2d(3, 2, 1), _stage(64, STAGES[0]))
        self.resnet_layer_3 = nn.Sequential(_stage(64, STAGES[1]))
        self.layer4 = _stage(128, STAGES[2])
        self.layer5 = _stage(256, STAGES[3])

    def forward(self, img):
        x_128 = self.resnet_layer_1(img)
        x_64 = self.resnet_layer_2(x_128)
        x_32 = self.resnet_layer_3(x_64)
        return self.layer5(self.layer4(x_32)), x_128, x_64


def _up(cin, c=256):
    mods = [nn.ConvTranspose2d(cin, c, 3, 2, 1, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True)]
    for _ in range(2):
        mods += [nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True)]
    return nn.Sequential(*mods)


class Decoder(nn.Module):
    """ASPP + the two x2 decoders of model/aspp.py:5-114.  forward() stops in front of conv_1x1_4."""

    def __init__(self, n_out=17):
        super().__init__()
        self.conv_1x1_1 = nn.Conv2d(512, 256, 1)
        self.bn_conv_1x1_1 = nn.BatchNorm2d(256)
        for i, d in enumerate((6, 12, 18), 1):
            setattr(self, "conv_3x3_%d" % i, nn.Conv2d(512, 256, 3, 1, d, d))
            setattr(self, "bn_conv_3x3_%d" % i, nn.BatchNorm2d(256))
        self.conv_1x1_2 = nn.Conv2d(512, 256, 1)
        self.bn_conv_1x1_2 = nn.BatchNorm2d(256)
        self.conv_1x1_3 = nn.Conv2d(1280, 256, 1)
        self.bn_conv_1x1_3 = nn.BatchNorm2d(256)
        self.upsample_1 = _up(256)
        self.upsample_2 = _up(256 + 64)
        self.conv_1x1_4 = nn.Conv2d(256 + 64, n_out, 1)

    # cuDNN (9.x, B200) runs the dilation-18 convolution of the 32 x 32 map in 225 ms at 128 crops on channels_last bf16
    # input (9.5 ms on NCHW; dilations 6 and 12 take 0.25 ms).  Two ways round it, both measured by tools/prof_net_body.py:
    # `tap_dilation` (default on CUDA: that layer as nine 1x1 taps, everything stays channels_last) or `aspp_nchw` (the
    # 134 MB feature map is re-laid once and the ASPP block alone runs in NCHW: layout only, same values).
    aspp_nchw = False
    tap_dilation = 0            # dilations >= this (and > 0) go through tap_conv3x3 instead of cuDNN's dilated kernel
    _ASPP_CONVS = ("conv_1x1_1", "conv_3x3_1", "conv_3x3_2", "conv_3x3_3", "conv_1x1_2", "conv_1x1_3")

    def use_nchw_aspp(self, on=True):
        self.aspp_nchw = bool(on)
        fmt = torch.contiguous_format if on else torch.channels_last
        for n in self._ASPP_CONVS:
            w = getattr(self, n).weight
            w.data = w.data.contiguous(memory_format=fmt)
        return self

    def forward(self, feat, x_64):
        if self.aspp_nchw:
            feat = feat.contiguous()
        branches = [_cb(self.conv_1x1_1, self.bn_conv_1x1_1, feat)]
        for i in (1, 2, 3):
            conv, bn = getattr(self, "conv_3x3_%d" % i), getattr(self, "bn_conv_3x3_%d" % i)
            if 0 < self.tap_dilation <= conv.dilation[0]:
                branches.append(F.relu(bn(tap_conv3x3(conv, feat)), inplace=True))
            else:
                branches.append(_cb(conv, bn, feat))
        # image-level branch: bilinear interpolation of a 1x1 map is a broadcast
        pooled = _cb(self.conv_1x1_2, self.bn_conv_1x1_2, feat.mean((2, 3), keepdim=True))
        branches.append(pooled.expand(-1, -1, feat.shape[2], feat.shape[3]))
        y = _cb(self.conv_1x1_3, self.bn_conv_1x1_3, torch.cat(branches, 1))
        return self.upsample_2(torch.cat([self.upsample_1(y), x_64], 1))


class ZebraNetBody(nn.Module):
    """img [B,3,256,256] -> (x [B,256,128,128], x_128 [B,64,128,128]); `tail` = the weights the fused head takes."""

    def __init__(self, n_out=17):
        super().__init__()
        self.resnet = Backbone()
        self.aspp = Decoder(n_out)

    def forward(self, img):
        feat, x_128, x_64 = self.resnet(img)
        return self.aspp(feat, x_64), x_128

    @property
    def tail(self):
        return self.aspp.conv_1x1_4

    def logits(self, img):
        """The unfused network output (model/aspp.py:112), for the comparison arm and the fixtures."""
        x, x_128 = self(img)
        return self.tail(torch.cat([x, x_128], 1))

    # ---- reference checkpoint format -------------------------------------------------------------------------------
    # model/resnet.py:184-196 registers the torchvision trunk twice (`resnet` and `resnet_layer_{1,2,3}` share modules),
    # so a reference state dict holds those tensors under two names; BinaryCodeNet_Deeplab adds the prefix `net.`.
    _ALIAS = (("resnet.resnet_layer_1.0.", "resnet.resnet.0."), ("resnet.resnet_layer_1.1.", "resnet.resnet.1."),
              ("resnet.resnet_layer_2.1.", "resnet.resnet.4."), ("resnet.resnet_layer_3.0.", "resnet.resnet.5."))

    def reference_state_dict(self, prefix="net."):
        out = {}
        for k, v in self.state_dict().items():
            out[prefix + k] = v
            for mine, dup in self._ALIAS:
                if k.startswith(mine):
                    out[prefix + dup + k[len(mine):]] = v
        return out

    def load_reference_state_dict(self, sd, prefix="net."):
        dups = tuple(prefix + d for _, d in self._ALIAS)
        own = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix) and not k.startswith(dups)}
        return self.load_state_dict(own, strict=True)

    # ---- inference form ----------------------------------------------------------------------------------------------
    def fold_batchnorm(self):
        """eval-mode BatchNorms folded into the preceding (transposed) convolution; the BatchNorm becomes Identity."""
        from torch.nn.utils.fusion import fuse_conv_bn_eval
        assert not self.training
        for parent in list(self.modules()):
            kids = list(parent.named_children())
            for (n0, m0), (n1, m1) in zip(kids, kids[1:]):
                if isinstance(m1, nn.BatchNorm2d) and isinstance(m0, (nn.Conv2d, nn.ConvTranspose2d)):
                    setattr(parent, n0, fuse_conv_bn_eval(m0, m1, transpose=isinstance(m0, nn.ConvTranspose2d)))
                    setattr(parent, n1, nn.Identity())
        assert not any(isinstance(m, nn.BatchNorm2d) for m in self.modules())
        return self


def build(seed=0, n_out=17, device="cpu", dtype=torch.float32, fold=False, aspp_nchw=False, tap_dilation=None):
    """The random-init network of configs[4]: torch.manual_seed(seed), default initialisers, eval."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = ZebraNetBody(n_out).eval()
    # a freshly constructed BatchNorm is the identity in eval mode (mean 0, var 1, weight 1, bias 0): give every one
    # seeded non-trivial statistics and affine terms so that folding them (and the checkpoint round trip) is a real test
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.05)
            m.running_var.copy_(1.0 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.6 + 0.2 * torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.05)
    torch.random.set_rng_state(gen_state)
    if fold:
        net.fold_batchnorm()
    net = net.to(device=device, dtype=dtype)
    if dtype in (torch.bfloat16, torch.float16) or str(device).startswith("cuda"):
        net = net.to(memory_format=torch.channels_last)
        if tap_dilation is None and not aspp_nchw:
            tap_dilation = 18 if str(device).startswith("cuda") else 0
    net.aspp.tap_dilation = int(tap_dilation or 0)
    if aspp_nchw:
        net.aspp.use_nchw_aspp(True)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def images(B, seed=0, size=256, device="cpu", dtype=torch.float32):
    """inputs N(0,1) [B,3,size,size], seeded per call on the CPU generator so every device sees the same."""
    g = torch.Generator().manual_seed(1005 * 65536 + seed)
    x = torch.randn(B, 3, size, size, generator=g)
    x = x.to(device=device, dtype=dtype)
    return x.contiguous(memory_format=torch.channels_last) if x.is_cuda or dtype != torch.float32 else x
