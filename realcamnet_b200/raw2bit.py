"""Host-side mirror of the reference's ``models/raw2bit.py`` (hot-path classes).

  CALayer 238-253, ResidualBlockWithCA 257-289, ConvTransBlock_mzj 292-328,
  HyCondMod{Conv,Enc,Dec}Block 730-814, HybridConditionModule 817-858,
  SpatialFeatureTransform 860-886, raw_compression_tcm_final 1614-2027 (THE paper model).
  GMABlock / GMAAtten / ConvGMABlock (168-184, 209-234, 330-355) wrap the GroupMix block.

Same constructor arguments, parameter names and forward()/compress()/decompress()/update()
signatures as the reference; every tensor op runs in librcn_b200.so.  Inference (eval) only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .entropy_models import CompressionModel, EntropyBottleneck, GaussianConditional, RansDecoder, rans_encode
from .groupmix import GMA_Block  # noqa: F401  (re-exported like the reference's copy, raw2bit.py:98-142)
from .layers import (AttentionBlock, Conv2d, Linear, ResidualBlock, ResidualBlockUpsample, ResidualBlockWithStride,
                     conv, conv1x1, conv3x3, subpel_conv3x3)
from .LiteISP import Color_Condition_GFM, Lens_Shading_Correction, Res_GFM
from .ops import (ACT_CLAMP01, ACT_GELU, ACT_HALF_TANH, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, EPI_MUL_AUXP1,
                  EPI_MULP1_AUX, STORE_NCHW, STORE_PS2_NCHW)
from .tcm import (Block, ConvTransBlock, SliceCodecModel, StageRunner, SWAtten, SwinBlock, _cc_transform,  # noqa: F401
                  _run_cc, get_scale_table)


class CALayer(nn.Module):
    """models/raw2bit.py:238-253 (bias-free squeeze/excite over Linear layers)."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def _f(self, x, res=None, out=None):
        g = ops.channel_mean(x)
        g = self.fc[2]._f(self.fc[0]._f(g, act=ACT_RELU), act=ACT_SIGMOID)
        return ops.scale_add(x, g.reshape(-1), per_n=True, res=res, out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class ResidualBlockWithCA(nn.Module):
    """models/raw2bit.py:257-289."""

    def __init__(self, in_ch: int, out_ch: int, redution=8):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.ca = CALayer(out_ch, redution)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def _f(self, x, out=None, presplit=None):
        t, tsp = self.conv1._f(x, act=ACT_LRELU, slope=0.01, presplit=presplit, emit_split=True, keep_fp32=False)
        t = self.conv2._f(t, presplit=tsp)
        identity = x if self.skip is None else self.skip._f(x, presplit=presplit)
        return self.ca._f(t, res=identity, out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class SpatialFeatureTransform(nn.Module):
    """models/raw2bit.py:860-886: x*scale(cond) + shift(cond) (+ x)."""

    def __init__(self, cond_channels=64, n_features=64, ada_method='vanilla', residual=True):
        super().__init__()
        assert ada_method == 'vanilla'
        self.cond_scale = nn.Sequential(Conv2d(cond_channels, n_features, 3, stride=1, padding=1), nn.ReLU(inplace=True),
                                        Conv2d(n_features, n_features, 3, stride=1, padding=1))
        self.cond_shift = nn.Sequential(Conv2d(cond_channels, n_features, 3, stride=1, padding=1), nn.ReLU(inplace=True),
                                        Conv2d(n_features, n_features, 3, stride=1, padding=1))
        self.residual = residual

    def _merged_first(self):
        """cond_scale[0] and cond_shift[0] as one packed conv (weights / biases concatenated on the output axis), re-packed when
        either changes."""
        a, b = self.cond_scale[0], self.cond_shift[0]
        if a.bias is None or b.bias is None or a.weight.shape != b.weight.shape:
            return None
        key = tuple((t._version, t.data_ptr()) for t in (a.weight, a.bias, b.weight, b.bias))
        hit = self.__dict__.get("_rcn_merged")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                pc = ops.pack_weight(torch.cat([a.weight.detach(), b.weight.detach()], 0).contiguous(),
                                     torch.cat([a.bias.detach(), b.bias.detach()], 0).contiguous())
            hit = self.__dict__["_rcn_merged"] = (key, pc)
        return hit[1]

    def _f(self, x, cond, extra=None, out=None, cond_split=None, split_out=None, keep_fp32=True):
        """x*scale + shift + x (+ extra); both element-wise steps are conv epilogues.
        cond_split: operand planes of cond (shared by every block of a level); split_out / keep_fp32: write the result as
        (also / only) the operand planes of the layer that reads it."""
        if not self.residual:
            raise NotImplementedError("residual=False is never used by the reference")
        sp = cond_split if cond_split is not None else \
            ops.shared_split(cond, [ops.pack(self.cond_scale[0]), ops.pack(self.cond_shift[0])])
        pcm = self._merged_first() if ops.bf16_planes_enabled() else None
        if pcm is not None and pcm.cout == 128:
            # both first layers read cond: ONE 3x3 conv with the two weight sets stacked on the output axis (N = 128 fills the MMA
            # width, the cond planes are fetched once); the second layers read their half of its planes
            _, both = ops.conv2d(cond, pcm, act=ACT_RELU, presplit=sp, emit_split=True, keep_fp32=False)
            s = h = None
            ssp, hsp = both.channels(0, 64), both.channels(64, 128)
        else:
            s, ssp = self.cond_scale[0]._f(cond, act=ACT_RELU, presplit=sp, emit_split=True, keep_fp32=False)
            h, hsp = self.cond_shift[0]._f(cond, act=ACT_RELU, presplit=sp, emit_split=True, keep_fp32=False)
        t = self.cond_scale[2]._f(s, epi=EPI_MULP1_AUX, aux=x, res=extra, presplit=ssp)      # (scale + 1) * x (+ extra)
        if split_out is None:
            return self.cond_shift[2]._f(h, res=t, out=out, presplit=hsp)                    # shift + ...
        return self.cond_shift[2]._f(h, res=t, out=out, presplit=hsp, split_out=split_out, keep_fp32=keep_fp32)[0]

    def forward(self, x, cond):
        return ops.to_nchw(self._f(ops.to_nhwc(x), ops.to_nhwc(cond)))


class ConvTransBlock_mzj(nn.Module):
    """Conv (CA + local RAW feature transform) || Swin block, models/raw2bit.py:292-328.
    forward([x, cond]) -> (x, cond)."""

    def __init__(self, conv_dim, trans_dim, head_dim, window_size, drop_path, type='W'):
        super().__init__()
        assert type in ['W', 'SW']
        self.conv_dim, self.trans_dim, self.head_dim = conv_dim, trans_dim, head_dim
        self.num_head = trans_dim // head_dim
        self.window_size, self.drop_path, self.type = window_size, drop_path, type
        self.trans_block = Block(trans_dim, trans_dim, head_dim, window_size, drop_path, type)
        self.conv1_1 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv1_2 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv_block = ResidualBlockWithCA(conv_dim, conv_dim, 8)
        self.spatial_transform = SpatialFeatureTransform(cond_channels=conv_dim, n_features=conv_dim)

    def _f(self, x, cond, out=None, cond_split=None, emit_stride=0, presplit=None):
        """emit_stride=2: returns (fea | None, polyphase operand planes) for a stride-2 consumer that is the only reader;
        emit_stride=1: returns (fea, operand planes of fea) for the next block's conv1_1; presplit: operand planes of x."""
        cd, td = self.conv_dim, self.trans_dim
        planes = ops.bf16_planes_enabled() and ops.plane_channels(cd) == cd and ops.plane_channels(td) == td and \
            ops.plane_channels(cd + td) == cd + td
        if not planes:
            both = self.conv1_1._f(x, presplit=presplit)
            cat = torch.empty_like(both)
            conv_identity = both[..., :cd]
            cx = self.conv_block._f(conv_identity)
            self.spatial_transform._f(cx, cond, extra=conv_identity, out=cat[..., :cd], cond_split=cond_split)
            self.trans_block._f(both[..., cd:], out=cat[..., cd:])
            fea = self.conv1_2._f(cat, res=x, out=out)
            return (fea, None) if emit_stride else fea
        # tcgen05 engine: see ConvTransBlock._f -- the concat read by conv1_2 only exists as operand planes
        both, bsp = self.conv1_1._f(x, emit_split=True, presplit=presplit)
        N, H, W, _ = both.shape
        csp = ops.alloc_planes(N, H, W, cd + td, both.device)
        conv_identity = both[..., :cd]
        cx = self.conv_block._f(conv_identity, presplit=bsp.channels(0, cd))
        self.spatial_transform._f(cx, cond, extra=conv_identity, cond_split=cond_split, split_out=csp.channels(0, cd), keep_fp32=False)
        self.trans_block._f(both[..., cd:], split_out=csp.channels(cd, cd + td), keep_fp32=False)
        if emit_stride == 2:
            return self.conv1_2._f(None, res=x, out=out, presplit=csp, emit_split=True, keep_fp32=False, emit_stride=2)
        if emit_stride == 1:
            return self.conv1_2._f(None, res=x, out=out, presplit=csp, emit_split=True)
        return self.conv1_2._f(None, res=x, out=out, presplit=csp)

    def forward(self, xx):
        x, cond = xx[0], xx[1]
        return ops.to_nchw(self._f(ops.to_nhwc(x), ops.to_nhwc(cond))), cond


class HyCondModConvBlock(nn.Module):
    """models/raw2bit.py:730-744."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, act='relu'):
        super().__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        if act == 'lrelu':
            self.act, self._act = nn.LeakyReLU(0.2, inplace=True), (ACT_LRELU, 0.2)
        elif act == 'prelu':
            raise NotImplementedError("prelu is never selected by the reference")
        else:
            self.act, self._act = nn.ReLU(inplace=True), (ACT_RELU, 0.0)

    def _f(self, x, out=None):
        return self.conv._f(x, act=self._act[0], slope=self._act[1], out=out)


class HyCondModEncBlock(nn.Module):
    """models/raw2bit.py:746-768 (stride downscale only, the reference default)."""

    def __init__(self, in_channels, out_channels, downscale_method='stride'):
        super().__init__()
        assert downscale_method == 'stride'
        self.down = HyCondModConvBlock(in_channels, out_channels, stride=2)
        self.conv = HyCondModConvBlock(out_channels, out_channels)

    def _f(self, x, out=None, split_out=None):
        """split_out: also write the result as operand planes (its half of a decoder concat's planes)"""
        t, sp = self.down.conv._f(x, act=self.down._act[0], slope=self.down._act[1], emit_split=True, keep_fp32=False)
        if split_out is not None:
            return self.conv.conv._f(t, act=self.conv._act[0], slope=self.conv._act[1], out=out, presplit=sp, split_out=split_out,
                                     keep_fp32=True)[0]
        return self.conv.conv._f(t, act=self.conv._act[0], slope=self.conv._act[1], out=out, presplit=sp)


class HyCondModDecBlock(nn.Module):
    """models/raw2bit.py:782-814 (bilinear x2, align_corners=True)."""

    def __init__(self, in_channels, out_channels, upscale_method='bilinear'):
        super().__init__()
        assert upscale_method == 'bilinear'
        self.up = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                                HyCondModConvBlock(in_channels, out_channels))
        self.conv = HyCondModConvBlock(in_channels, out_channels)

    def _f(self, x1, cat, out=None, cat_planes=None):
        """cat: (N,2H,2W,2*out) buffer whose first half already holds the skip tensor x2.
        cat_planes: operand planes of the concat whose first half the skip's producer has written: the up-sampled map, the up
        conv's half of the concat and the concat itself then only exist as operand planes (no fp32 maps, no rcn_split_bf16 passes)."""
        oc = self.up[1].conv.out_channels
        if cat_planes is not None:
            u, usp = ops.upsample_bilinear2x(x1, emit_split=True)
            up = self.up[1]
            up.conv._f(u, act=up._act[0], slope=up._act[1], presplit=usp, split_out=cat_planes.channels(oc, 2 * oc), keep_fp32=False)
            return self.conv.conv._f(None, act=self.conv._act[0], slope=self.conv._act[1], out=out, presplit=cat_planes)
        self.up[1]._f(ops.upsample_bilinear2x(x1), out=cat[..., oc:])
        return self.conv._f(cat, out=out)


class HybridConditionModule(nn.Module):
    """Local RAW condition UNet, models/raw2bit.py:817-858.  forward(x) -> [cond@1/2, cond@1/4, cond@1/8]."""

    def __init__(self, in_channels=4, out_channels=64, init_mid_channels=16, down_method='stride', up_method='bilinear'):
        super().__init__()
        m = init_mid_channels
        self.in_conv = HyCondModConvBlock(in_channels, m)
        self.enc_1 = HyCondModEncBlock(m, m * 2, down_method)
        self.enc_2 = HyCondModEncBlock(m * 2, m * 4, down_method)
        self.enc_3 = HyCondModEncBlock(m * 4, m * 8, down_method)
        self.dec_1 = HyCondModDecBlock(m * 8, m * 4, up_method)
        self.dec_2 = HyCondModDecBlock(m * 4, m * 2, up_method)
        self.dec_3 = HyCondModDecBlock(m * 2, m, up_method)
        self.out_conv = HyCondModConvBlock(m, out_channels)
        oc = out_channels
        self.CondNet1 = nn.Sequential(Conv2d(oc, oc, 3, 2, 1), nn.LeakyReLU(0.1, True), Conv2d(oc, oc, 1))
        self.CondNet2 = nn.Sequential(Conv2d(oc, oc, 3, 2, 1), nn.LeakyReLU(0.1, True), Conv2d(oc, oc, 3, 2, 1))
        self.CondNet3 = nn.Sequential(Conv2d(oc, oc, 3, 2, 1), nn.LeakyReLU(0.1, True), Conv2d(oc, oc, 3, 2, 1),
                                      nn.LeakyReLU(0.1, True), Conv2d(oc, oc, 3, 2, 1))
        self._m = m

    def _f(self, x, presplit=None):
        N, H, W, _ = x.shape
        m = self._m
        # skip tensors are produced straight into the first half of the decoder concat buffers
        planes = ops.bf16_planes_enabled() and ops.get_engine() == "bf16x3" and all(ops.plane_channels(c) == c for c in (m, 2 * m, 4 * m, 8 * m))
        if planes:
            # tcgen05 engine: the decoder concats only exist as operand planes, written half by half by the skip's producer and by the
            # up conv (whose own input comes as planes from the up-sampling kernel); the skips keep an fp32 copy for the encoder
            p3 = ops.alloc_planes(N, H, W, 2 * m, x.device)
            p2 = ops.alloc_planes(N, H // 2, W // 2, 4 * m, x.device)
            p1 = ops.alloc_planes(N, H // 4, W // 4, 8 * m, x.device)
            x1 = self.in_conv.conv._f(x, act=self.in_conv._act[0], slope=self.in_conv._act[1], presplit=presplit,
                                      split_out=p3.channels(0, m), keep_fp32=True)[0]
            x2 = self.enc_1._f(x1, split_out=p2.channels(0, 2 * m))
            x3 = self.enc_2._f(x2, split_out=p1.channels(0, 4 * m))
            x4 = self.enc_3._f(x3)
            y = self.dec_1._f(x4, None, cat_planes=p1)
            y = self.dec_2._f(y, None, cat_planes=p2)
            y = self.dec_3._f(y, None, cat_planes=p3)
        else:
            cat3 = ops.empty(N, H, W, 2 * m, like=x)
            cat2 = ops.empty(N, H // 2, W // 2, 4 * m, like=x)
            cat1 = ops.empty(N, H // 4, W // 4, 8 * m, like=x)
            x1 = self.in_conv.conv._f(x, act=self.in_conv._act[0], slope=self.in_conv._act[1], out=cat3[..., :m], presplit=presplit)
            x2 = self.enc_1._f(x1, out=cat2[..., :2 * m])
            x3 = self.enc_2._f(x2, out=cat1[..., :4 * m])
            x4 = self.enc_3._f(x3)
            y = self.dec_1._f(x4, cat1)
            y = self.dec_2._f(y, cat2)
            y = self.dec_3._f(y, cat3)
        # the three CondNets read y through a stride-2 3x3 conv only: out_conv writes their (shared) polyphase operand planes itself
        y, sp = self.out_conv.conv._f(y, act=self.out_conv._act[0], slope=self.out_conv._act[1], emit_split=True, keep_fp32=False,
                                      emit_stride=2)
        if sp is None:
            heads = [self.CondNet1[0], self.CondNet2[0], self.CondNet3[0]]
            sp = ops.shared_split(y, [ops.pack(h) for h in heads], stride=2)
        t, tsp = self.CondNet1[0]._f(y, act=ACT_LRELU, slope=0.1, presplit=sp, emit_split=True, keep_fp32=False)
        c1 = self.CondNet1[2]._f(t, presplit=tsp)
        # stride-2 -> stride-2 chains: each layer writes the next one's polyphase operand planes itself
        t2, sp2 = self.CondNet2[0]._f(y, act=ACT_LRELU, slope=0.1, presplit=sp, emit_split=True, keep_fp32=False, emit_stride=2)
        c2 = self.CondNet2[2]._f(t2, presplit=sp2)
        t3, sp3 = self.CondNet3[0]._f(y, act=ACT_LRELU, slope=0.1, presplit=sp, emit_split=True, keep_fp32=False, emit_stride=2)
        t3, sp3 = self.CondNet3[2]._f(t3, act=ACT_LRELU, slope=0.1, presplit=sp3, emit_split=True, keep_fp32=False, emit_stride=2)
        c3 = self.CondNet3[4]._f(t3, presplit=sp3)
        return [c1, c2, c3]

    def forward(self, x):
        return [ops.to_nchw(c) for c in self._f(ops.to_nhwc(x))]


class RBU(nn.Module):
    """models/raw2bit.py:3181-3206 (ResidualBlockUpsample without IGDN)."""

    def __init__(self, in_ch: int, out_ch: int, upsample: int = 2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def _f(self, x, out=None):
        sp = ops.shared_split(x, [ops.pack(self.subpel_conv[0]), ops.pack(self.upsample[0])])
        t, tsp = self.subpel_conv._f(x, act=ACT_LRELU, slope=0.01, presplit=sp, emit_split=True, keep_fp32=False)
        t = self.conv._f(t, presplit=tsp)
        return self.upsample._f(x, res=t, out=out, presplit=sp)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class raw_compression_tcm_final(SliceCodecModel):
    """The paper model (models/raw2bit.py:1614-2027).

    forward(x=[raw(B,4,H,W), cond(B,4,h',w'), coord(B,2,H,W)]) -> dict(x_hat, y, lft, lsc, likelihoods{y,z},
    para{means,scales,y});  compress(x) -> {"strings": [[y_bytes], [z_bytes]*B], "shape"};
    decompress(strings, shape) -> {"x_hat"};  update(scale_table=None, force=False).
    """

    def __init__(self, config=[2, 2, 2, 2, 2, 2, 2], head_dim=[8, 16, 32, 32, 16, 8, 8], drop_path_rate=0, N=64, M=320,
                 num_slices=5, max_support_slices=5, **kwargs):
        super().__init__()
        if drop_path_rate:
            raise NotImplementedError("inference path: drop_path_rate must be 0")
        self.config, self.head_dim, self.window_size = config, head_dim, 8
        self.num_slices, self.max_support_slices = num_slices, max_support_slices
        dim, self.M, self.N = N, M, N
        ws = self.window_size
        cond_c, modulation_blocks = 128, 1
        self.classifier = Color_Condition_GFM(in_channels=4, out_c=cond_c)
        self.lsc = Lens_Shading_Correction(in_channels=2, out_c=2 * N, nf=2 * N)
        self.local_condition = HybridConditionModule(out_channels=N, init_mid_channels=16)
        self.conv_first = conv3x3(4, 2 * N)
        self.conv_down = ResidualBlockWithStride(2 * N, 2 * N, 2)

        def typ(i):
            return 'W' if not i % 2 else 'SW'

        def gfm():
            return nn.Sequential(*[Res_GFM(in_nc=2 * N, chan=2 * N, cond_c=cond_c, nf=4 * N) for _ in range(modulation_blocks)])

        self.gfm1 = gfm()
        self.m_down1 = nn.Sequential(*[ConvTransBlock_mzj(dim, dim, head_dim[0], ws, 0, typ(i)) for i in range(config[0])])
        self.m_down1_down = ResidualBlockWithStride(2 * N, 2 * N, stride=2)
        self.gfm2 = gfm()
        self.m_down2 = nn.Sequential(*[ConvTransBlock_mzj(dim, dim, head_dim[1], ws, 0, typ(i)) for i in range(config[1])])
        self.m_down2_down = ResidualBlockWithStride(2 * N, 2 * N, stride=2)
        self.gfm3 = gfm()
        self.m_down3 = nn.Sequential(*[ConvTransBlock_mzj(dim, dim, head_dim[2], ws, 0, typ(i)) for i in range(config[2])])
        self.m_down3_down = conv3x3(2 * N, M, stride=2)

        m_up1 = [ConvTransBlock(dim, dim, head_dim[3], ws, 0, typ(i)) for i in range(config[3])] + [ResidualBlockUpsample(2 * N, 2 * N, 2)]
        m_up2 = [ConvTransBlock(dim, dim, head_dim[4], ws, 0, typ(i)) for i in range(config[4])] + [ResidualBlockUpsample(2 * N, 2 * N, 2)]
        m_up3 = [ConvTransBlock(dim, dim, head_dim[5], ws, 0, typ(i)) for i in range(config[5])] + [subpel_conv3x3(2 * N, 2 * N, 2)]
        tail = [ResidualBlock(2 * N, 2 * N), subpel_conv3x3(2 * N, 3, 2)]
        self.g_s = nn.Sequential(*[ResidualBlockUpsample(M, 2 * N, 2)] + m_up1 + m_up2 + m_up3 + tail)

        self.h_a = nn.Sequential(*[ResidualBlockWithStride(320, 2 * N, 2)] +
                                 [ConvTransBlock(N, N, 32, 4, 0, typ(i)) for i in range(config[0])] + [conv3x3(2 * N, 192, stride=2)])
        self.h_mean_s = nn.Sequential(*[ResidualBlockUpsample(192, 2 * N, 2)] +
                                      [ConvTransBlock(N, N, 32, 4, 0, typ(i)) for i in range(config[3])] + [subpel_conv3x3(2 * N, 320, 2)])
        self.h_scale_s = nn.Sequential(*[ResidualBlockUpsample(192, 2 * N, 2)] +
                                       [ConvTransBlock(N, N, 32, 4, 0, typ(i)) for i in range(config[3])] + [subpel_conv3x3(2 * N, 320, 2)])

        sl = 320 // num_slices
        self.atten_mean = nn.ModuleList(nn.Sequential(SWAtten(320 + sl * min(i, 5), 320 + sl * min(i, 5), 16, ws, 0, inter_dim=128))
                                        for i in range(num_slices))
        self.atten_scale = nn.ModuleList(nn.Sequential(SWAtten(320 + sl * min(i, 5), 320 + sl * min(i, 5), 16, ws, 0, inter_dim=128))
                                         for i in range(num_slices))
        self.cc_mean_transforms = nn.ModuleList(_cc_transform(320 + sl * min(i, 5), sl) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_cc_transform(320 + sl * min(i, 5), sl) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_cc_transform(320 + sl * min(i + 1, 6), sl) for i in range(num_slices))
        self.entropy_bottleneck = EntropyBottleneck(192)
        self.gaussian_conditional = GaussianConditional(None)

    # ------------------------------------------------------------------------------ transforms (NHWC)
    def _analysis(self, x):
        """models/raw2bit.py:1771-1796 (and 1877-1901 in compress)."""
        raw, cond = ops.to_nhwc(x[0]), ops.to_nhwc(x[1])
        vec = self.classifier._f(cond)                                  # (B,1,1,128) gfm_vector
        rsp = ops.shared_split(raw, [ops.pack(self.conv_first), ops.pack(self.local_condition.in_conv.conv)])   # both read raw
        local = self.local_condition._f(raw, presplit=rsp)
        if ops.fused_ingest_ok(self.lsc.layers(), x[2], self.conv_first) and tuple(x[2].shape[2:]) == tuple(raw.shape[1:3]):
            # fused ingest (csrc/ingest.cu): lens-shading MLP + conv_first(x) * (lsc + 1) in one kernel; lsc is written once as the
            # NCHW map forward() returns (raw2bit.py:1853), the product once as the polyphase operand planes of conv_down
            lsc_fea, fsp = self.lsc._f_fused(x[2], raw, self.conv_first, emit_stride=2)
            fea = None
        else:
            lsc_fea = self.lsc._f(ops.to_nhwc(x[2]), nchw=True)
            # conv_first(x) * (lsc + 1): only conv_down (stride 2) reads it -> written once, as polyphase operand planes
            fea, fsp = self.conv_first._f(raw, epi=EPI_MUL_AUXP1, aux=lsc_fea, aux_nchw=True, presplit=rsp, emit_split=True,
                                          keep_fp32=False, emit_stride=2)
        # every layer below hands its result to the next one as operand planes written by its own epilogue (the fp32 map is kept
        # only where a residual / aux operand needs it): no rcn_split_bf16 pass between the layers of a level
        fea, gsp = self.conv_down._f(fea, presplit=fsp, emit_split=True)
        for lvl, (gfm, blocks, down) in enumerate(((self.gfm1, self.m_down1, self.m_down1_down),
                                                   (self.gfm2, self.m_down2, self.m_down2_down),
                                                   (self.gfm3, self.m_down3, self.m_down3_down))):
            for g in gfm:
                fea, gsp = g._f(fea, vec, presplit=gsp, emit_split=True)
            csp = ops.shared_split(local[lvl], [ops.pack(blocks[0].spatial_transform.cond_scale[0]),
                                                ops.pack(blocks[0].spatial_transform.cond_shift[0])])   # cond planes: once per level
            fsp = None
            for j, blk in enumerate(blocks):
                if j == len(blocks) - 1:     # the level's last block feeds the stride-2 `down` layer only
                    fea, fsp = blk._f(fea, local[lvl], cond_split=csp, emit_stride=2, presplit=gsp)
                else:
                    fea, gsp = blk._f(fea, local[lvl], cond_split=csp, emit_stride=1, presplit=gsp)
            if lvl < 2:
                fea, gsp = down._f(fea, presplit=fsp, emit_split=True)
            else:
                fea = down._f(fea, presplit=fsp)
        return fea, lsc_fea, local

    def _g_s(self, y_hat, clamp=False):
        h = y_hat
        mods = list(self.g_s)
        ts = self.tail_start if self.tail_start is not None else len(mods) - 3
        sp = None
        for i, m in enumerate(mods[:-3]):
            with self._tail_scope(i >= ts):     # experiments only (profiles/r2_precision_policy.md): a tail starting earlier
                # layer -> layer hand-over as operand planes (no split pass), except across an engine boundary and into the tail
                chain = isinstance(m, (ConvTransBlock, ResidualBlockUpsample)) and i < ts
                nxt_ok = chain and (i + 1 < len(mods) - 3) and (i + 1 < ts) and isinstance(mods[i + 1], (ConvTransBlock, ResidualBlockUpsample))
                if chain:
                    r = m._f(h, presplit=sp, emit_split=nxt_ok)
                    h, sp = r if nxt_ok else (r, None)
                else:
                    h, sp = m._f(h), None
        # tail at full resolution (raw2bit.py:1680-1682): subpel -> ResidualBlock -> subpel.  The 128-channel maps are 2.1 GB
        # each: the producers write the consumers' bf16 operand planes themselves, and the ResidualBlock output (read by the
        # last conv only) never exists in fp32.
        # Precision policy (SliceCodecModel.tail_engine): these four convs run as one fp16 pass.
        with self._tail_scope():
            h, hsp = mods[-3]._f(h, emit_split=True, keep_fp32=True)
            h, rsp = mods[-2]._f(h, presplit=hsp, emit_split=True, keep_fp32=False)    # h is None when only the planes exist
            return mods[-1]._f(h, presplit=rsp, store=STORE_PS2_NCHW, act=ACT_CLAMP01 if clamp else ACT_NONE)

    # ------------------------------------------------------------------------------ public API
    @torch.no_grad()
    def forward(self, x, emit_strings=False):
        if self.training:
            raise NotImplementedError("inference path only: call .eval() (train mode adds quantisation noise)")

        def stage_a(xs):
            y, lsc_fea, local = self._analysis(xs)
            E = self._entropy_stage(y, emit_strings)
            E.y, E.lsc_fea, E.lft = y, lsc_fea, local[2]
            return E

        def stage_b(E):
            # the coder front end (CDF lookups) ran inside the Gaussian kernel of stage a; its D2H copy is started between the
            # stages so the host state chain runs WHILE the synthesis transform g_s (half of the FLOPs) executes
            y_nchw = ops.to_nchw(E.y)
            return {"x_hat": self._g_s(E.ms[..., 320:]), "y": y_nchw, "lft": ops.to_nchw(E.lft), "lsc": E.lsc_fea,
                    "likelihoods": {"y": ops.to_nchw(E.y_lik), "z": ops.to_nchw(E.z_lik)},
                    "para": {"means": ops.to_nchw(E.means), "scales": ops.to_nchw(E.scales), "y": y_nchw}}

        run = StageRunner(self, ("forward", emit_strings), list(x), stage_a, stage_b)
        E = run.a()
        pending = self._begin_host_copy(E.coder.packed, E.coder.raw, E.coder.flags, E.z_sym) if emit_strings else None
        out = dict(run.b())
        if emit_strings:
            out["strings"] = self._strings(E, pending)
            out["shape"] = torch.Size(E.z.shape[1:3])
        return out

    def _compress_stage(self, xs):
        return self._entropy_stage(self._analysis(xs)[0], True, want_lik=False)

    @torch.no_grad()
    def compress(self, x):
        """models/raw2bit.py:1876-1960."""
        if self.gaussian_conditional._offset.numel() == 0:
            raise RuntimeError("call update() before compress()")
        run = StageRunner(self, ("compress",), list(x), self._compress_stage, None)
        E = run.a()
        pending = self._begin_host_copy(E.coder.packed, E.coder.raw, E.coder.flags, E.z_sym)
        return {"strings": self._strings(E, pending), "shape": torch.Size(E.z.shape[1:3])}

    @torch.no_grad()
    def decompress(self, strings, shape):
        """models/raw2bit.py:1982-2027 (batch 1, like the reference)."""
        return {"x_hat": self._g_s(self._decode_stage(strings, shape), clamp=True)}


# ----------------------------------------------------------------------------- GroupMix drop-ins
class GMABlock(nn.Module):
    """Two GMA_Blocks on an NCHW map (models/raw2bit.py:168-184): the GroupMix replacement of SwinBlock."""

    def __init__(self, input_dim, head_dim, drop_path) -> None:
        super().__init__()
        self.input_dim = input_dim
        self.num_head = input_dim // head_dim
        self.block_1 = GMA_Block(input_dim, self.num_head, drop_path_rate=drop_path)
        self.block_2 = GMA_Block(input_dim, self.num_head, drop_path_rate=drop_path)

    def _f(self, x):
        return self.block_2._f(self.block_1._f(x))

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class GMAAtten(AttentionBlock):
    """GroupMix-gated attention of the entropy parameter nets (models/raw2bit.py:209-234; SWAtten with GMABlock inside)."""

    def __init__(self, input_dim, output_dim, head_dim, drop_path, inter_dim=192) -> None:
        if inter_dim is None:
            # the reference's inter_dim=None branch builds GMABlock but then calls the undefined in_conv/out_conv
            # (models/raw2bit.py:219-234): it can never run
            raise NotImplementedError("GMAAtten(inter_dim=None) cannot run in the reference (forward needs in_conv/out_conv)")
        super().__init__(N=inter_dim)
        self.num_head = inter_dim // head_dim
        self.head_dim = head_dim
        self.non_local_block = GMABlock(inter_dim, head_dim, drop_path=drop_path)
        self.in_conv = conv1x1(input_dim, inter_dim)
        self.out_conv = conv1x1(inter_dim, output_dim)

    def _f(self, x, out=None):
        x = self.in_conv._f(x)
        z = self.non_local_block._f(x)
        return self.out_conv._f(self._gate(x, z, x), out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))


class ConvGMABlock(nn.Module):
    """Conv || GroupMix block, shape-compatible with ConvTransBlock (models/raw2bit.py:330-355)."""

    def __init__(self, conv_dim, trans_dim, head_dim, drop_path=0.):
        super().__init__()
        self.conv_dim, self.trans_dim, self.head_dim = conv_dim, trans_dim, head_dim
        self.num_head = trans_dim // head_dim
        self.drop_path = drop_path
        self.trans_block = GMA_Block(trans_dim, self.num_head, drop_path_rate=drop_path)
        self.conv1_1 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv1_2 = Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1, 1, 0, bias=True)
        self.conv_block = ResidualBlock(conv_dim, conv_dim)

    def _f(self, x, out=None):
        cd = self.conv_dim
        both = self.conv1_1._f(x)
        cat = torch.empty_like(both)
        self.conv_block._f(both[..., :cd], out=cat[..., :cd], extra_identity=True)
        self.trans_block._f(both[..., cd:], out=cat[..., cd:])
        return self.conv1_2._f(cat, res=x, out=out)

    def forward(self, x):
        return ops.to_nchw(self._f(ops.to_nhwc(x)))
