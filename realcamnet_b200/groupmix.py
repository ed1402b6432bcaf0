"""Host-side mirror of the reference's ``models/groupmix.py`` GroupMix attention block (hot-path classes).

  Mlp 21-38, Agg_0 41-53, Aggregator 56-105, ConvRelPosEnc 108-156, EfficientAtt 159-200,
  ConvPosEnc 203-217, SeparableConv2d 240-249, GMA_Block 274-299 (copy at models/raw2bit.py:98-142).
Token tensors are (B, N=H*W, C) row-major exactly like the reference (N index = h*W + w), which is the
NHWC layout of the kernels, so forward(x, size) needs no transposes.  Inference (eval) only: the
SyncBatchNorm layers are applied with their running statistics.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _C, ops
from .layers import Conv2d, Linear
from .ops import ACT_GELU, ACT_HSWISH, ACT_NONE


def _taps(conv: nn.Conv2d):
    """depthwise weight (C,1,k,k) -> [k*k][C] (cached on the module)."""
    w = conv.weight
    key = (w.data_ptr(), w._version)
    hit = getattr(conv, "_rcn_taps", None)
    if hit is None or hit[0] != key:
        C, k = w.shape[0], w.shape[-1]
        hit = (key, w.detach().reshape(C, k * k).t().contiguous())
        conv._rcn_taps = hit
    return hit[1]


def _bn_fold(bn):
    """eval-mode BatchNorm as y = x*g + b."""
    key = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
    hit = getattr(bn, "_rcn_fold", None)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            b = bn.bias - bn.running_mean * g
        hit = (key, g.contiguous(), b.contiguous())
        bn._rcn_fold = hit
    return hit[1], hit[2]


class Mlp(nn.Module):
    """models/groupmix.py:21-38."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        assert act_layer is nn.GELU and drop == 0.
        self.fc1 = Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def _f(self, x, res=None, out=None):
        h, sp = self.fc1._f(x, act=ACT_GELU, emit_split=True, keep_fp32=False)
        return self.fc2._f(h, res=res, out=out, presplit=sp)

    def forward(self, x):
        shp = x.shape
        return self._f(x.reshape(1, 1, -1, shp[-1]).contiguous()).reshape(shp[:-1] + (-1,))


class SeparableConv2d(nn.Module):
    """models/groupmix.py:240-249: depthwise k x k then pointwise 1x1 (bias-free by default)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=False):
        super().__init__()
        assert stride == 1 and dilation == 1 and padding == kernel_size // 2
        self.conv1 = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, dilation, groups=in_channels, bias=bias)
        self.pointwise_conv = Conv2d(in_channels, out_channels, 1, 1, 0, 1, 1, bias=bias)

    def _f(self, x, **kw):
        k = self.conv1.kernel_size[0]
        d = ops.depthwise_conv(x, _taps(self.conv1), self.conv1.bias, k)
        return self.pointwise_conv._f(d, **kw)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class Agg_0(nn.Module):
    """models/groupmix.py:41-53."""

    def __init__(self, seg_dim):
        super().__init__()
        self.conv = SeparableConv2d(seg_dim * 3, seg_dim, 3, 1, 1)
        self.norm = nn.LayerNorm(seg_dim)
        self.act = nn.Hardswish()


class Aggregator(nn.Module):
    """models/groupmix.py:56-105 (multi-scale depthwise aggregation of q, k, v)."""

    def __init__(self, dim, seg=4):
        super().__init__()
        self.dim, self.seg = dim, seg
        seg_dim = self.dim // self.seg
        self.norm0 = nn.SyncBatchNorm(seg_dim)
        self.act0 = nn.Hardswish()
        self.agg1 = SeparableConv2d(seg_dim, seg_dim, 3, 1, 1)
        self.norm1 = nn.SyncBatchNorm(seg_dim)
        self.act1 = nn.Hardswish()
        self.agg2 = SeparableConv2d(seg_dim, seg_dim, 5, 1, 2)
        self.norm2 = nn.SyncBatchNorm(seg_dim)
        self.act2 = nn.Hardswish()
        self.agg3 = SeparableConv2d(seg_dim, seg_dim, 7, 1, 3)
        self.norm3 = nn.SyncBatchNorm(seg_dim)
        self.act3 = nn.Hardswish()
        self.agg0 = Agg_0(seg_dim)

    def _pw_folded(self, j):
        """pointwise conv of agg{j} with eval BatchNorm norm{j} folded in (cached)."""
        agg, bn = getattr(self, f"agg{j}"), getattr(self, f"norm{j}")
        w = agg.pointwise_conv.weight
        g, b = _bn_fold(bn)
        key = (w.data_ptr(), w._version, g.data_ptr())
        hit = getattr(agg, "_rcn_pw", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, ops.pack_weight((w.detach() * g[:, None, None, None]).contiguous(), b))
            agg._rcn_pw = hit
        return hit[1]

    def _f(self, qkv3, B):
        """qkv3: (3B,H,W,C) -- q images, then k images, then v images (the reference's (3B,N,C) view).
        Returns (agg (3B,H,W,4C/5) head-major, x_local (B,H,W,C/5))."""
        B3, H, W, C = qkv3.shape
        s = self.dim // self.seg
        agg = ops.empty(B3, H, W, 4 * s, like=qkv3)
        g0, b0 = _bn_fold(self.norm0)
        ops.scale_add(qkv3[..., :s], g0, b0, per_n=False, act=ACT_HSWISH, out=agg[..., :s])
        for j in (1, 2, 3):
            sep = getattr(self, f"agg{j}")
            d = ops.depthwise_conv(qkv3[..., j * s:(j + 1) * s], _taps(sep.conv1), None, sep.conv1.kernel_size[0])
            ops.conv2d(d, self._pw_folded(j), act=ACT_HSWISH, out=agg[..., j * s:(j + 1) * s])
        # local branch: channel-concat of segment 4 of q, k, v (groupmix.py:93)
        loc_in = ops.empty(B, H, W, 3 * s, like=qkv3)
        taps = _taps(self.agg0.conv.conv1)
        for t in range(3):
            ops.depthwise_conv(qkv3[t * B:(t + 1) * B, :, :, 4 * s:5 * s], taps[:, t * s:(t + 1) * s].contiguous(), None, 3,
                               out=loc_in[..., t * s:(t + 1) * s])
        loc = self.agg0.conv.pointwise_conv._f(loc_in)
        return agg, loc

    def forward(self, x, size, num_head):
        B3, Ntok, C = x.shape
        H, W = size
        assert Ntok == H * W
        agg, loc = self._f(x.reshape(B3, H, W, C).contiguous(), B3 // 3)
        loc = ops.layernorm(loc, self.agg0.norm.weight, self.agg0.norm.bias, self.agg0.norm.eps, act=ACT_HSWISH)
        Ct = agg.shape[-1]
        x_out = agg.reshape(3, B3 // 3, Ntok, num_head, Ct // num_head).permute(0, 1, 3, 2, 4)
        return x_out, loc.reshape(B3 // 3, Ntok, -1)


class ConvRelPosEnc(nn.Module):
    """models/groupmix.py:108-156: depthwise conv over v per head group, times q."""

    def __init__(self, Ch, h, window):
        super().__init__()
        if isinstance(window, int):
            window = {window: h}
        elif not isinstance(window, dict):
            raise ValueError()
        self.window = window
        self.conv_list = nn.ModuleList()
        self.head_splits = []
        for cur_window, cur_head_split in window.items():
            pad = cur_window // 2
            self.conv_list.append(nn.Conv2d(cur_head_split * Ch, cur_head_split * Ch, kernel_size=(cur_window, cur_window),
                                            padding=(pad, pad), dilation=(1, 1), groups=cur_head_split * Ch))
            self.head_splits.append(cur_head_split)
        self.channel_splits = [x * Ch for x in self.head_splits]

    def _f(self, q, v):
        """q, v: (B,H,W,Ct) head-major -> q * conv(v)"""
        out = torch.empty((q.shape[0], q.shape[1], q.shape[2], q.shape[3]), device=q.device, dtype=torch.float32)
        c0 = 0
        for conv, cj in zip(self.conv_list, self.channel_splits):
            ops.depthwise_conv(v[..., c0:c0 + cj], _taps(conv), conv.bias, conv.kernel_size[0], mul=q[..., c0:c0 + cj],
                               out=out[..., c0:c0 + cj])
            c0 += cj
        return out

    def forward(self, q, v, size):
        B, h, Ntok, Ch = q.shape
        H, W = size
        assert Ntok == H * W
        qi = q.permute(0, 2, 1, 3).reshape(B, H, W, h * Ch).contiguous()
        vi = v.permute(0, 2, 1, 3).reshape(B, H, W, h * Ch).contiguous()
        return self._f(qi, vi).reshape(B, Ntok, h, Ch).permute(0, 2, 1, 3)


class EfficientAtt(nn.Module):
    """models/groupmix.py:159-200.  forward(x(B,N,C), size=(H,W)) -> (B,N,C)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.aggregator = Aggregator(dim=dim, seg=5)
        trans_dim = dim // 5 * 4
        self.crpe = ConvRelPosEnc(Ch=trans_dim // num_heads, h=num_heads, window={3: 2, 5: 3, 7: 3})
        self.dim = dim

    def _qkv_parts(self):
        w, b = self.qkv.weight, self.qkv.bias
        key = (w.data_ptr(), w._version, None if b is None else b._version)
        hit = getattr(self, "_rcn_qkv", None)
        if hit is None or hit[0] != key:
            C = self.dim
            parts = [ops.pack_weight(w.detach()[t * C:(t + 1) * C].contiguous(),
                                     None if b is None else b.detach()[t * C:(t + 1) * C].contiguous()) for t in range(3)]
            hit = (key, parts)
            self._rcn_qkv = hit
        return hit[1]

    def _f(self, x, res=None, out=None):
        """x: (B,H,W,C) token map"""
        B, H, W, C = x.shape
        s = C // 5
        Ct, heads = 4 * s, self.num_heads
        Ch = Ct // heads
        qkv3 = ops.empty(3 * B, H, W, C, like=x)
        for t, pc in enumerate(self._qkv_parts()):
            ops.conv2d(x, pc, out=qkv3[t * B:(t + 1) * B])
        agg, loc = self.aggregator._f(qkv3, B)
        q, k, v = agg[:B], agg[B:2 * B], agg[2 * B:]
        crpe = self.crpe._f(q, v)
        cat = ops.empty(B, H, W, C, like=x)
        n0 = self.aggregator.agg0.norm
        ops.layernorm(loc, n0.weight, n0.bias, n0.eps, out=cat[..., Ct:], act=ACT_HSWISH)
        lib = _C.lib()
        wsf = int(lib.rcn_groupmix_workspace_floats(B, H * W, heads, Ch))
        ws = torch.empty((wsf,), device=x.device, dtype=torch.float32)
        kv = torch.empty((B, heads, Ch, Ch), device=x.device, dtype=torch.float32)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        ld = lambda t: ops.geom(t)[4]
        att_out = cat[..., :Ct]
        _C.check(lib.rcn_groupmix_attention(P(q), ld(q), P(k), ld(k), P(v), ld(v), P(crpe), ld(crpe), B, H * W, heads, Ch,
                                            float(self.scale), P(att_out), ld(att_out), P(kv), P(ws), wsf, ops._stream()),
                 "rcn_groupmix_attention")
        return self.proj._f(cat, res=res, out=out)

    def forward(self, x, size):
        B, Ntok, C = x.shape
        H, W = size
        assert Ntok == H * W
        return self._f(x.reshape(B, H, W, C).contiguous()).reshape(B, Ntok, C)


class ConvPosEnc(nn.Module):
    """models/groupmix.py:203-217: depthwise 3x3 + identity on the token map."""

    def __init__(self, dim, k=3):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, k, 1, k // 2, groups=dim)

    def _f(self, x):
        return ops.depthwise_conv(x, _taps(self.proj), self.proj.bias, self.proj.kernel_size[0], add_input=True)

    def forward(self, x, size):
        B, Ntok, C = x.shape
        H, W = size
        assert Ntok == H * W
        return self._f(x.reshape(B, H, W, C).contiguous()).reshape(B, Ntok, C)


class GMA_Block(nn.Module):
    """models/groupmix.py:274-299.  forward(x(B,N,C), size) -> (B,N,C)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path_rate=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop_path_rate or drop or attn_drop:
            raise NotImplementedError("inference path: dropout rates must be 0 (reference defaults)")
        self.cpe = ConvPosEnc(dim=dim, k=3)
        self.norm1 = norm_layer(dim)
        self.att = EfficientAtt(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path_rate = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _f(self, x, out=None):
        x = self.cpe._f(x)
        cur = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x = self.att._f(cur, res=x)
        cur = ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp._f(cur, res=x, out=out)

    def forward(self, x_input, size):
        B, Ntok, C = x_input.shape
        H, W = size
        assert Ntok == H * W
        return self._f(x_input.reshape(B, H, W, C).contiguous()).reshape(B, Ntok, C)
