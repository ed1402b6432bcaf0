"""Host-side mirror of the reference's ``models/networks.py`` block library (hot-path subset).

  seq            models/networks.py:117-128     conv (string-mode factory)  models/networks.py:146-221
  DWTForward/DWTInverse  models/networks.py:224-249     DWTForward_/DWTInverse_  models/networks.py:10-48
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
                while j < len(mods) and isinstance(mods[j], (nn.ReLU, nn.LeakyReLU, nn.PReLU, nn.PixelShuffle)):
                    n = mods[j]
                    if isinstance(n, nn.ReLU) and act == ACT_NONE:
                        act = ACT_RELU
                    elif isinstance(n, nn.LeakyReLU) and act == ACT_NONE:
                        act, slope = ACT_LRELU, n.negative_slope
                    elif isinstance(n, nn.PReLU) and act == ACT_NONE and store == STORE_NHWC:
                        act, slope = ACT_LRELU, prelu_slope(n)
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


_prelu_cache = {}


def prelu_slope(m: nn.PReLU) -> float:
    """nn.PReLU() with its single shared parameter (the only form on the path, models/LiteISP.py:2160-2193) is LeakyReLU with a
    learned slope: read once per parameter version (a device->host read, outside any graph capture)."""
    if m.weight.numel() != 1:
        raise NotImplementedError("per-channel PReLU is not on the B200 path")
    key = (m.weight._version, m.weight.data_ptr())
    hit = _prelu_cache.get(id(m))
    if hit is None or hit[0] != key or hit[2]() is not m.weight:
        import weakref
        hit = _prelu_cache[id(m)] = (key, float(m.weight.detach().reshape(-1)[0].item()), weakref.ref(m.weight))
    return hit[1]


class Down2x2(nn.Conv2d):
    """nn.Conv2d(C, O, 2, 2): the learned down-sampler of ISPUNet_GFM_LSC / ResUNet (models/LiteISP.py:1253,2056).  Runs as
    rcn_space_to_depth2 + a 1x1 contraction over the (i, j, c)-ordered weights."""

    def __init__(self, in_channels, out_channels):
        super().__init__(in_channels, out_channels, 2, 2)
        self._packed = None

    def _pack(self):
        w, b = self.weight, self.bias
        key = (w._version, w.data_ptr(), None if b is None else (b._version, b.data_ptr()))
        if self._packed is None or self._packed[0] != key or self._packed[2] is not w:
            O, C = w.shape[0], w.shape[1]
            w2 = w.detach().permute(0, 2, 3, 1).reshape(O, 4 * C).contiguous()     # [o][(i*2+j)*C + c]
            self._packed = (key, ops.pack_weight(w2, b), w)
        return self._packed[1]

    def _f(self, x, **kw):
        return ops.conv2d(ops.space_to_depth2(x), self._pack(), **kw)

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


class DWTForward_(_Haar):
    """models/networks.py:10-27: the functional twin of DWTForward (the flipped filter bank it builds equals DWTForward's, one
    (4,1,2,2) weight shared by all channels)."""

    def __init__(self):
        super().__init__()
        w = torch.tensor([[[[0.5, 0.5], [0.5, 0.5]]], [[[0.5, 0.5], [-0.5, -0.5]]],
                          [[[0.5, -0.5], [0.5, -0.5]]], [[[0.5, -0.5], [-0.5, 0.5]]]])
        self.weight = nn.Parameter(w, requires_grad=False)

    def _f(self, x):
        return ops.dwt_forward(x)


class DWTInverse_(_Haar):
    """models/networks.py:30-48."""

    def __init__(self):
        super().__init__()
        w = torch.tensor([[[[0.5, 0.5], [0.5, 0.5]]], [[[0.5, 0.5], [-0.5, -0.5]]],
                          [[[0.5, -0.5], [0.5, -0.5]]], [[[0.5, -0.5], [-0.5, 0.5]]]])
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
