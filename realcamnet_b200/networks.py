"""Host-side mirror of the reference's ``models/networks.py`` block library (hot-path subset).

  seq            models/networks.py:117-128     conv (string-mode factory)  models/networks.py:146-221
  DWTForward/DWTInverse  models/networks.py:224-249
  CALayer 255-270, RCABlock 296-311, RCAGroup 317-335
Only the layer modes that occur on the RAW->sRGB path are accepted by ``conv``: 'C', 'R'/'r', 'L'/'l', '2'.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops
from .layers import Conv2d
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, STORE_NHWC, STORE_PS2


def seq(*args):
    """models/networks.py:117-128 -- collapses a single module, otherwise nn.Sequential."""
    if len(args) == 1:
        args = args[0]
    if isinstance(args, nn.Module):
        return args
    modules = OrderedDict()
    if isinstance(args, OrderedDict):
        for k, v in args.items():
            modules[k] = seq(v)
        return _Seq(modules)
    assert isinstance(args, (list, tuple))
    return _Seq(*[seq(i) for i in args])


class _Seq(nn.Sequential):
    """nn.Sequential whose NHWC path fuses Conv2d + ReLU/LeakyReLU + PixelShuffle(2) runs into one launch."""

    def _f(self, x, res=None):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            last_conv = isinstance(m, Conv2d) and not any(isinstance(n, Conv2d) or hasattr(n, "_f") for n in mods[i + 1:])
            if isinstance(m, Conv2d):
                act, slope, store, j = ACT_NONE, 0.0, STORE_NHWC, i + 1
                while j < len(mods) and isinstance(mods[j], (nn.ReLU, nn.LeakyReLU, nn.PixelShuffle)):
                    n = mods[j]
                    if isinstance(n, nn.ReLU) and act == ACT_NONE:
                        act = ACT_RELU
                    elif isinstance(n, nn.LeakyReLU) and act == ACT_NONE:
                        act, slope = ACT_LRELU, n.negative_slope
                    elif isinstance(n, nn.PixelShuffle) and store == STORE_NHWC and n.upscale_factor == 2:
                        store = STORE_PS2
                    else:
                        break
                    j += 1
                x = m._f(x, act=act, slope=slope, store=store, res=res if (last_conv and j >= len(mods)) else None)
                if last_conv and j >= len(mods):
                    res = None
                i = j
            elif hasattr(m, "_f"):
                x = m._f(x)
                i += 1
            else:
                raise NotImplementedError(f"layer {type(m).__name__} is not on the B200 path")
        if res is not None:
            raise RuntimeError("residual could not be fused")
        return x

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


def conv(in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, output_padding=0, dilation=1, groups=1,
         bias=True, padding_mode='zeros', mode='CBR'):
    """models/networks.py:146-221 (modes used on the path)."""
    L = []
    for t in mode:
        if t == 'C':
            assert groups == 1 and dilation == 1 and padding == kernel_size // 2 and padding_mode == 'zeros'
            L.append(Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias))
        elif t == 'R':
            L.append(nn.ReLU(inplace=True))
        elif t == 'r':
            L.append(nn.ReLU(inplace=False))
        elif t == 'L':
            L.append(nn.LeakyReLU(negative_slope=1e-1, inplace=True))
        elif t == 'l':
            L.append(nn.LeakyReLU(negative_slope=1e-1, inplace=False))
        elif t == '2':
            L.append(nn.PixelShuffle(upscale_factor=2))
        else:
            raise NotImplementedError('Undefined type on the B200 path: {}'.format(t))
    return seq(*L)


class _Haar(nn.Module):
    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class DWTForward(_Haar):
    """models/networks.py:224-235; the fixed +-0.5 filter bank is kept as the ``weight`` tensor for
    state_dict compatibility, the kernel hard-codes it."""

    def __init__(self, in_channels=64):
        super().__init__()
        w = torch.tensor([[[[0.5, 0.5], [0.5, 0.5]]], [[[0.5, 0.5], [-0.5, -0.5]]],
                          [[[0.5, -0.5], [0.5, -0.5]]], [[[0.5, -0.5], [-0.5, 0.5]]]]).repeat(in_channels, 1, 1, 1)
        self.weight = nn.Parameter(w, requires_grad=False)

    def _f(self, x):
        return ops.dwt_forward(x)


class DWTInverse(_Haar):
    """models/networks.py:238-249."""

    def __init__(self, in_channels=64):
        super().__init__()
        w = torch.tensor([[[[0.5, 0.5], [0.5, 0.5]]], [[[0.5, 0.5], [-0.5, -0.5]]],
                          [[[0.5, -0.5], [0.5, -0.5]]], [[[0.5, -0.5], [-0.5, 0.5]]]]).repeat(in_channels // 4, 1, 1, 1)
        self.weight = nn.Parameter(w, requires_grad=False)

    def _f(self, x):
        return ops.dwt_inverse(x)


class CALayer(nn.Module):
    """Channel attention: x * sigmoid(W2 relu(W1 gap(x)))  (models/networks.py:255-270)."""

    def __init__(self, channel=64, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_du = nn.Sequential(Conv2d(channel, channel // reduction, 1, padding=0, bias=True), nn.ReLU(inplace=True),
                                     Conv2d(channel // reduction, channel, 1, padding=0, bias=True), nn.Sigmoid())

    def _gate(self, x):
        g = ops.channel_mean(x)
        g = self.conv_du[0]._f(g, act=ACT_RELU)
        return self.conv_du[2]._f(g, act=ACT_SIGMOID)

    def _f(self, x, res=None, out=None):
        return ops.scale_add(x, self._gate(x).reshape(-1), per_n=True, res=res, out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class RCABlock(nn.Module):
    """models/networks.py:296-311."""

    def __init__(self, in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True, mode='CRC', reduction=16):
        super().__init__()
        assert in_channels == out_channels
        if mode[0] in ['R', 'L']:
            mode = mode[0].lower() + mode[1:]
        self.res = conv(in_channels, out_channels, kernel_size, stride, padding, bias=bias, mode=mode)
        self.ca = CALayer(out_channels, reduction)

    def _f(self, x):
        return self.ca._f(self.res._f(x), res=x)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class RCAGroup(nn.Module):
    """models/networks.py:317-335."""

    def __init__(self, in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True, mode='CRC', reduction=16, nb=12):
        super().__init__()
        assert in_channels == out_channels
        if mode[0] in ['R', 'L']:
            mode = mode[0].lower() + mode[1:]
        RG = [RCABlock(in_channels, out_channels, kernel_size, stride, padding, bias, mode, reduction) for _ in range(nb)]
        RG.append(conv(out_channels, out_channels, mode='C'))
        self.rg = nn.Sequential(*RG)

    def _f(self, x):
        h = x
        n = len(self.rg)
        for i in range(n - 1):
            h = self.rg[i]._f(h)
        return self.rg[n - 1]._f(h, res=x)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))
