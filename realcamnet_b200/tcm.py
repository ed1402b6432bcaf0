"""Host-side mirror of the reference's ``models/tcm.py`` block library (hot-path classes).

Class names, constructor arguments, parameter names and call signatures follow the reference:
  WMSA            models/tcm.py:139-212      Block          models/tcm.py:214-236
  ConvTransBlock  models/tcm.py:242-268      SWAtten        models/tcm.py:270-291
  SwinBlock       models/tcm.py:293-312      get_scale_table models/tcm.py:33-34
All tensor math is dispatched to librcn_b200.so through ``ops``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .layers import (ACT_GELU, AttentionBlock, Conv2d, Linear, ResidualBlock, conv, conv1x1, conv3x3,  # noqa: F401
                     subpel_conv3x3)

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """models/tcm.py:33-34 (host-side, CPU tensor: the table must be bit-identical to the reference's)."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class WMSA(nn.Module):
    """Window multi-head self-attention of the Swin blocks (models/tcm.py:139-212).

    forward(x[b,h,w,c]) -> [b,h,w,c].  embedding_layer and linear run as 1x1 contractions
    (rcn_conv2d), the attention core as rcn_wmsa (bias gather + shift mask computed from indices).
    """

    def __init__(self, input_dim, output_dim, head_dim, window_size, type):
        super().__init__()
        self.input_dim, self.output_dim, self.head_dim = input_dim, output_dim, head_dim
        self.scale = head_dim ** -0.5
        self.n_heads = input_dim // head_dim
        self.window_size = window_size
        self.type = type
        self.embedding_layer = Linear(input_dim, 3 * input_dim, bias=True)
        p = torch.zeros((2 * window_size - 1) * (2 * window_size - 1), self.n_heads)
        nn.init.trunc_normal_(p, std=.02)
        # stored as (heads, 2w-1, 2w-1) like the reference after its view/transposes (models/tcm.py:158)
        self.relative_position_params = nn.Parameter(
            p.view(2 * window_size - 1, 2 * window_size - 1, self.n_heads).transpose(1, 2).transpose(0, 1).contiguous())
        self.linear = Linear(input_dim, output_dim)

    def _f(self, x, res=None, out=None):
        qkv = self.embedding_layer._f(x)
        rel = self.relative_position_params
        if not rel.is_contiguous():
            rel = rel.contiguous()
        o = ops.wmsa(qkv, rel.detach(), self.head_dim, self.window_size, self.type != 'W')
        return self.linear._f(o, res=res, out=out)

    def forward(self, x):
        return self._f(x.contiguous())


class Block(nn.Module):
    """Swin block: x + WMSA(LN(x)); x + MLP(LN(x))  (models/tcm.py:214-236). NHWC in/out."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, type='W', input_resolution=None):
        super().__init__()
        assert type in ['W', 'SW']
        if drop_path:
            raise NotImplementedError("inference path: drop_path must be 0 (the reference default)")
        self.input_dim, self.output_dim, self.type = input_dim, output_dim, type
        self.ln1 = nn.LayerNorm(input_dim)
        self.msa = WMSA(input_dim, input_dim, head_dim, window_size, self.type)
        self.drop_path = nn.Identity()
        self.ln2 = nn.LayerNorm(input_dim)
        self.mlp = nn.Sequential(Linear(input_dim, 4 * input_dim), nn.GELU(), Linear(4 * input_dim, output_dim))

    def _f(self, x, out=None):
        t = ops.layernorm(x, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        x1 = self.msa._f(t, res=x)
        t = ops.layernorm(x1, self.ln2.weight, self.ln2.bias, self.ln2.eps)
        h, hsp = self.mlp[0]._f(t, act=ACT_GELU, emit_split=True, keep_fp32=False)   # 4C-wide hidden: planes only
        return self.mlp[2]._f(h, res=x1, out=out, presplit=hsp)

    def forward(self, x):
        return self._f(x.contiguous())


class ConvTransBlock(nn.Module):
    """Parallel conv / Swin-transformer block (models/tcm.py:242-268). NCHW forward, NHWC ``_f``."""

    def __init__(self, conv_dim, trans_dim, head_dim, window_size, drop_path, type='W'):
        super().__init__()
        assert type in ['W', 'SW']
        self.conv_dim, self.trans_dim, self.head_dim = conv_dim, trans_dim, head_dim
        self.window_size, self.drop_path, self.type = window_size, drop_path, type
        self.trans_block = Block(trans_dim, trans_dim, head_dim, window_size, drop_path, type)
        self.conv1_1 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv1_2 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv_block = ResidualBlock(conv_dim, conv_dim)

    def _f(self, x, out=None):
        cd = self.conv_dim
        both = self.conv1_1._f(x)                      # torch.split -> channel views
        cat = torch.empty_like(both)
        self.conv_block._f(both[..., :cd], out=cat[..., :cd], extra_identity=True)
        self.trans_block._f(both[..., cd:], out=cat[..., cd:])
        return self.conv1_2._f(cat, res=x, out=out)     # x + conv1_2(cat(conv_x, trans_x))

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class SwinBlock(nn.Module):
    """W-block followed by SW-block (models/tcm.py:293-312)."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path) -> None:
        super().__init__()
        self.block_1 = Block(input_dim, output_dim, head_dim, window_size, drop_path, type='W')
        self.block_2 = Block(input_dim, output_dim, head_dim, window_size, drop_path, type='SW')
        self.window_size = window_size

    def _f(self, x):
        N, H, W, C = x.shape
        if W <= self.window_size or H <= self.window_size:
            # the reference pads such maps and never crops them back (models/tcm.py:301-312), which then
            # fails inside WMSA's window rearrange; same condition -> same kind of error here
            raise ValueError(f"SwinBlock: {H}x{W} map must be larger than the window {self.window_size}")
        return self.block_2._f(self.block_1._f(x))

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class SWAtten(AttentionBlock):
    """Swin-attention gate of the entropy parameter nets (models/tcm.py:270-291)."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, inter_dim=192) -> None:
        if inter_dim is not None:
            super().__init__(N=inter_dim)
            self.non_local_block = SwinBlock(inter_dim, inter_dim, head_dim, window_size, drop_path)
            self.in_conv = conv1x1(input_dim, inter_dim)
            self.out_conv = conv1x1(inter_dim, output_dim)
        else:
            super().__init__(N=input_dim)
            self.non_local_block = SwinBlock(input_dim, input_dim, head_dim, window_size, drop_path)
        self._has_io = inter_dim is not None

    def _f(self, x, out=None):
        if self._has_io:
            x = self.in_conv._f(x)
        z = self.non_local_block._f(x)
        if self._has_io:
            g = self._gate(x, z, x)
            return self.out_conv._f(g, out=out)
        return self._gate(x, z, x, out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


def ste_round(x):
    """models/tcm.py:36-37; forward value only (inference path)."""
    raise NotImplementedError("use the fused entropy kernels (ops.eb_forward / ops.gaussian_conditional)")
