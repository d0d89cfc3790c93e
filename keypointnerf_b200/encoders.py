"""2-D image encoders that feed the ray-march path, emitting the kernel-native layout directly.

``HGFilterV2`` (one stacked hourglass, group norm; reference ``src/utils.py:322-414`` with ``ConvBlock`` 416-474,
``HourGlass`` 262-309, ``DeconvReLUGroup`` 311-320) produces the two geometry feature maps (64 channels at 1/8 and
8 channels at 1/2 of the encoder input); ``ResBlkEncoder`` (reference ``src/utils.py:216-259``, ``ResBlk`` 199-214)
the 8-channel texture map at 1/4.  The module / parameter NAMES are the reference's, so that its checkpoints load with
``strict=True`` (``load_ckpt``, reference ``src/model.py:113-117``); the convolutions themselves are library calls
(cuDNN through torch, channels-last) -- they are not part of the hand-written hot path.

What is specific to this build:
  * the whole stack runs in ``torch.channels_last``; the outputs are returned as NCHW-shaped tensors whose MEMORY is
    ``[V][H][W][C]`` -- exactly the atlas layout the CUDA kernels gather from, so ``kpn_set_scene`` takes them without
    the re-layout pass (``kpn_scene.layout = KPN_LAYOUT_NHWC``);
  * ``FeatureCache`` keeps the maps of the last source-image set: the reference re-runs both encoders for every rendered
    camera (``src/model.py:913-914`` after 479); a 90-camera sweep runs them once.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(min(32, ch), ch)


class ConvBlock(nn.Module):
    """Pre-activation residual block whose three 3x3 convolutions (out/2, out/4, out/4 channels) are concatenated."""

    def __init__(self, in_planes: int, out_planes: int, norm: str = "group"):
        super().__init__()
        if norm != "group":
            raise NotImplementedError("only the group-norm encoder of configs/zju.json is built")
        widths = (out_planes // 2, out_planes // 4, out_planes // 4)
        ins = (in_planes, widths[0], widths[1])
        for i, (ci, co) in enumerate(zip(ins, widths), 1):
            setattr(self, f"bn{i}", _gn(ci))
            setattr(self, f"conv{i}", nn.Conv2d(ci, co, 3, 1, 1, bias=False))
        self.bn4 = _gn(in_planes)
        self.downsample = None
        if in_planes != out_planes:
            self.downsample = nn.Sequential(self.bn4, nn.ReLU(True), nn.Conv2d(in_planes, out_planes, 1, bias=False))

    def forward(self, x):
        parts, y = [], x
        for i in (1, 2, 3):
            y = getattr(self, f"conv{i}")(F.relu(getattr(self, f"bn{i}")(y)))
            parts.append(y)
        skip = x if self.downsample is None else self.downsample(x)
        return torch.cat(parts, 1) + skip


class HourGlass(nn.Module):
    """Recursive hourglass of ``depth`` levels (Newell et al. 2016) with bicubic up-sampling (align_corners=True)."""

    def __init__(self, depth: int, num_features: int, norm: str = "group"):
        super().__init__()
        self.depth = depth
        for level in range(depth, 0, -1):
            for tag in ("b1_", "b2_", "b3_"):
                self.add_module(f"{tag}{level}", ConvBlock(num_features, num_features, norm))
        self.add_module("b2_plus_1", ConvBlock(num_features, num_features, norm))

    def _level(self, level: int, x):
        up = self._modules[f"b1_{level}"](x)
        low = self._modules[f"b2_{level}"](F.avg_pool2d(x, 2, stride=2))
        low = self._level(level - 1, low) if level > 1 else self._modules["b2_plus_1"](low)
        low = self._modules[f"b3_{level}"](low)
        return up + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)

    def forward(self, x):
        return self._level(self.depth, x)


class DeconvReLUGroup(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, bias: bool = False):
        super().__init__()
        self.conv = nn.ConvTranspose2d(in_ch, out_ch, 3, stride=2, padding=1, output_padding=1, bias=bias)
        self.nl = nn.ReLU(inplace=True)
        self.norm = _gn(out_ch)

    def forward(self, x):
        return self.nl(self.norm(self.conv(x)))


class HGFilterV2(nn.Module):
    def __init__(self, in_ch=3, out_ch=128, n_stack=2, n_downsample=4, norm="group", hd=False, **kwargs):
        super().__init__()
        if norm != "group":
            raise NotImplementedError("only the group-norm encoder of configs/zju.json is built")
        self.n_stack, self.hd = n_stack, hd
        self.unpack1 = DeconvReLUGroup(128, 32)
        self.conv_out = nn.Conv2d(32, kwargs.get("out_ch_hd", 8), 5, padding=2)
        self.conv1 = nn.Conv2d(in_ch, 64, 7, stride=2, padding=3)
        self.bn1 = nn.GroupNorm(32, 64)
        self.conv2, self.conv3, self.conv4 = ConvBlock(64, 128), ConvBlock(128, 128), ConvBlock(128, 256)
        for i in range(n_stack):
            self.add_module(f"m{i}", HourGlass(n_downsample, 256))
            self.add_module(f"top_m_{i}", ConvBlock(256, 256))
            self.add_module(f"conv_last{i}", nn.Conv2d(256, 256, 1))
            self.add_module(f"bn_end{i}", nn.GroupNorm(32, 256))
            self.add_module(f"l{i}", nn.Conv2d(256, out_ch, 1))
            if i < n_stack - 1:
                self.add_module(f"bl{i}", nn.Conv2d(256, 256, 1))
                self.add_module(f"al{i}", nn.Conv2d(out_ch, 256, 1))

    def forward(self, x):
        M = self._modules
        x = self.conv2(F.relu(self.bn1(self.conv1(x))))
        x_hd = self.conv_out(self.unpack1(x))
        if not self.hd:
            x = F.avg_pool2d(x, 2, stride=2)
        prev = self.conv4(self.conv3(x))
        out = None
        for i in range(self.n_stack):
            ll = M[f"top_m_{i}"](M[f"m{i}"](prev))
            ll = F.relu(M[f"bn_end{i}"](M[f"conv_last{i}"](ll)))
            out = M[f"l{i}"](ll)
            if i < self.n_stack - 1:
                prev = prev + M[f"bl{i}"](ll) + M[f"al{i}"](out)
        return [out, x_hd]


class ResBlk(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.layers = nn.Sequential(nn.ReplicationPad2d(1), nn.Conv2d(ch, ch, 3), nn.InstanceNorm2d(ch), nn.ReLU(True),
                                    nn.ReplicationPad2d(1), nn.Conv2d(ch, ch, 3), nn.InstanceNorm2d(ch))

    def forward(self, x):
        return x + self.layers(x)


class ResBlkEncoder(nn.Module):
    def __init__(self, in_ch=3, out_ch=8, ngf=16, n_downsample=3, n_blocks=4, n_upsample=3, norm="instance"):
        super().__init__()
        if norm != "instance":
            raise NotImplementedError("only the instance-norm texture encoder of configs/zju.json is built")
        IN, relu = nn.InstanceNorm2d, lambda: nn.ReLU(True)
        seq = [nn.ReplicationPad2d(3), nn.Conv2d(in_ch, ngf, 7), IN(ngf), relu()]
        ch = ngf
        for _ in range(n_downsample):
            seq += [nn.Conv2d(ch, 2 * ch, 3, stride=2, padding=1), IN(2 * ch), relu()]
            ch *= 2
        seq += [ResBlk(ch) for _ in range(n_blocks)]
        for _ in range(n_upsample):
            seq += [nn.ConvTranspose2d(ch, ch // 2, 3, stride=2, padding=1, output_padding=1), IN(ch // 2), relu()]
            ch //= 2
        if n_upsample > 0:
            seq += [nn.ReplicationPad2d(3), nn.Conv2d(ch, out_ch, 7)]
        self.layers = nn.Sequential(*seq)

    def forward(self, x):
        return self.layers(x)


def to_channels_last(module: nn.Module) -> nn.Module:
    """Weights of every convolution in channels-last memory: cuDNN then runs NHWC kernels end to end."""
    return module.to(memory_format=torch.channels_last)


def nhwc_view(t: torch.Tensor) -> torch.Tensor:
    """NCHW-shaped tensor whose storage is dense [N][H][W][C] (a no-op for the output of a channels-last network)."""
    return t.contiguous(memory_format=torch.channels_last)


def is_nhwc(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and t.shape[1] > 1


class FeatureCache:
    """Feature maps of the last source-image set, keyed by the image tensor's identity and version."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, im: torch.Tensor, fn):
        key = (im.data_ptr(), im._version, tuple(im.shape), str(im.device))
        if key != self.key:
            self.val = fn(im)
            self.key = key
            self._keep = im   # the key is an address: keep the tensor alive while its features are cached
        return self.val

    def clear(self):
        self.key, self.val, self._keep = None, None, None
