"""Host-side mirrors of the CompressAI layers the reference builds its codec from
(``from compressai.layers import ...``, models/tcm.py:4-11, models/raw2bit.py:11).

Same class names, constructor arguments and parameter/buffer names as compressai.layers, so a
reference ``state_dict`` loads unchanged; ``forward`` keeps the NCHW-in/NCHW-out contract while
``_f`` is the NHWC-internal path used when blocks are chained.  All arithmetic runs in
librcn_b200.so (no torch math on the hot path, no CPU fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .ops import (ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, EPI_GDN, EPI_IGDN, EPI_NONE, EPI_SIGMOID_GATE,
                  STORE_NHWC, STORE_PS2)


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameter holder (k in {1,3}, stride in {1,2}, padding k//2) running rcn_conv2d."""

    def _f(self, x, **kw):
        return ops.conv2d(x, ops.pack(self), stride=self.stride[0], **kw)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class Linear(nn.Linear):
    """nn.Linear on the last dim == 1x1 conv on an NHWC view."""

    def _f(self, x, **kw):
        return ops.conv2d(x, ops.pack(self), **kw)

    def forward(self, x):
        shp = x.shape
        y = self._f(x.reshape(1, 1, -1, shp[-1]).contiguous())
        return y.reshape(*shp[:-1], self.out_features)


def conv3x3(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def conv1x1(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """models/tcm.py:130-137 (only kernel_size=3 / stride=1 occurs on the path)."""
    return Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


class SubpelConv3x3(nn.Sequential):
    """compressai.layers.subpel_conv3x3 = Sequential(Conv2d, PixelShuffle(r)); the shuffle is fused
    into the conv's store (RCN_STORE_PS2)."""

    def __init__(self, in_ch, out_ch, r=2):
        assert r == 2, "only upsample=2 occurs in the reference"
        super().__init__(Conv2d(in_ch, out_ch * r * r, kernel_size=3, padding=1), nn.PixelShuffle(r))

    def _f(self, x, store=STORE_PS2, **kw):
        return self[0]._f(x, store=store, **kw)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


def subpel_conv3x3(in_ch, out_ch, r=1):
    return SubpelConv3x3(in_ch, out_ch, r)


class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def effective(self, p):
        """parameter-space map max(p, bound)^2 - pedestal (evaluated once per weight version)."""
        return torch.max(p, self.lower_bound.bound) ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalized divisive normalisation: x * rsqrt(beta + gamma . x^2) (inverse: * sqrt).
    One rcn_conv2d launch: 1x1 contraction over x^2 with the RCN_EPI_GDN / RCN_EPI_IGDN epilogue."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))
        self._packed = None

    def _pack(self):
        key = (self.beta.data_ptr(), self.beta._version, self.gamma.data_ptr(), self.gamma._version)
        if self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                beta = self.beta_reparam.effective(self.beta)
                gamma = self.gamma_reparam.effective(self.gamma)
            self._packed = (key, ops.pack_weight(gamma, beta))
        return self._packed[1]

    def _f(self, x, res=None, out=None, presplit=None, emit_split=False):
        """presplit: planes of x*x emitted by the producer of x (conv2d(..., emit_square=True)); emit_split: also write the result
        as the next layer's operand planes -> returns (out, planes)."""
        return ops.conv2d(x, self._pack(), in_square=True, epi=EPI_IGDN if self.inverse else EPI_GDN, aux=x, res=res,
                          out=out, presplit=presplit, emit_split=emit_split)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class _Block(nn.Module):
    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class ResidualBlockWithStride(_Block):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    def _f(self, x, out=None, presplit=None, emit_split=False):
        """presplit: operand planes of x emitted by its producer for this stride (x may then be None when a skip conv exists).
        emit_split: returns (out, planes of out) for the layer that reads the block's result."""
        # conv1 and the 1x1 skip read the same tensor with the same stride: one bf16 split serves both
        sp = presplit
        if sp is None and self.skip is not None:
            sp = ops.shared_split(x, [ops.pack(self.conv1), ops.pack(self.skip)], self.conv1.stride[0])
        t, tsp = self.conv1._f(x, act=ACT_LRELU, slope=0.01, presplit=sp, emit_split=True, keep_fp32=False)
        t, tsq = self.conv2._f(t, presplit=tsp, emit_split=True, emit_square=True)     # GDN's norm pool reads t*t: planes from here
        identity = x if self.skip is None else self.skip._f(x, presplit=sp)
        return self.gdn._f(t, res=identity, out=out, presplit=tsq, emit_split=emit_split)


class ResidualBlockUpsample(_Block):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def _f(self, x, out=None, presplit=None, emit_split=False):
        """presplit: operand planes of x from its producer; emit_split: returns (out, planes of out)."""
        sp = presplit if presplit is not None else ops.shared_split(x, [ops.pack(self.subpel_conv[0]), ops.pack(self.upsample[0])])
        t, tsp = self.subpel_conv._f(x, act=ACT_LRELU, slope=0.01, presplit=sp, emit_split=True, keep_fp32=False)
        t, tsq = self.conv._f(t, presplit=tsp, emit_split=True, emit_square=True)
        t = self.igdn._f(t, presplit=tsq)
        return self.upsample._f(x, res=t, out=out, presplit=sp, emit_split=emit_split)


class ResidualBlock(_Block):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def _f(self, x, out=None, extra_identity=False, presplit=None, emit_split=False, keep_fp32=True, split_out=None):
        """extra_identity: also add x once more (ConvTransBlock's ``conv_block(x) + x``).
        presplit / emit_split / keep_fp32: operand planes of x from its producer / planes of the result for its consumer."""
        t, tsp = self.conv1._f(x, act=ACT_LRELU, slope=0.01, emit_split=True, keep_fp32=False, presplit=presplit)
        identity = x if self.skip is None else self.skip._f(x, presplit=presplit)
        if extra_identity and self.skip is not None:
            raise ValueError("extra_identity needs in_ch == out_ch")
        return self.conv2._f(t, act=ACT_LRELU, slope=0.01, res=identity, res_scale=2.0 if extra_identity else 1.0, out=out,
                             presplit=tsp, emit_split=emit_split, keep_fp32=keep_fp32, split_out=split_out)


class AttentionBlock(_Block):
    """compressai.layers.AttentionBlock: a(x) * sigmoid(b(x)) + x with three ResidualUnits per branch."""

    def __init__(self, N):
        super().__init__()

        class ResidualUnit(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                          nn.ReLU(inplace=True), conv1x1(N // 2, N))
                self.relu = nn.ReLU(inplace=True)

            def _f(self, x, presplit=None, emit_split=False):
                """presplit: operand planes of x from its producer; emit_split: returns (y, planes of y | None) so that a chain of
                units (and the 1x1 layer after it) needs no rcn_split_bf16 pass between them."""
                t, sp = self.conv[0]._f(x, act=ACT_RELU, emit_split=True, keep_fp32=False, presplit=presplit)
                t, sp = self.conv[2]._f(t, act=ACT_RELU, presplit=sp, emit_split=True, keep_fp32=False)
                if emit_split:
                    return self.conv[4]._f(t, res=x, res_pre=True, act=ACT_RELU, presplit=sp, emit_split=True, keep_fp32=True)
                return self.conv[4]._f(t, res=x, res_pre=True, act=ACT_RELU, presplit=sp)

        self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())
        self.conv_b = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit(), conv1x1(N, N))

    def _gate(self, x, z, identity, out=None):
        """conv_a(x) * sigmoid(conv_b(z)) + identity"""
        a, b = x, z
        asp = bsp = None
        for i in range(3):      # unit -> unit -> 1x1 hand-over as operand planes (no rcn_split_bf16 pass in between)
            if i < 2:
                a, asp = self.conv_a[i]._f(a, presplit=asp, emit_split=True)
            else:
                a = self.conv_a[i]._f(a, presplit=asp)
            b, bsp = self.conv_b[i]._f(b, presplit=bsp, emit_split=True)
        return self.conv_b[3]._f(b, epi=EPI_SIGMOID_GATE, aux=a, res=identity, out=out, presplit=bsp)

    def _f(self, x):
        return self._gate(x, x, x)


__all__ = ["Conv2d", "Linear", "conv3x3", "conv1x1", "conv", "subpel_conv3x3", "SubpelConv3x3", "GDN",
           "ResidualBlock", "ResidualBlockWithStride", "ResidualBlockUpsample", "AttentionBlock",
           "ACT_NONE", "ACT_RELU", "ACT_LRELU", "ACT_GELU", "EPI_NONE", "STORE_NHWC"]
