"""Host-side mirror of the reference's ``models/LiteISP.py`` (hot-path classes).

  color_block 23-30, Color_Condition_GFM 345-361, Lens_Shading_Correction 363-378,
  Res_GFM 537-559, LiteISPNet_GFM_LSC 1924-2035 (BASELINE config 1).
forward([raw, cond, coord]) -> (B,3,2H,2W), as in the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import networks as N
from . import ops
from .layers import Conv2d, Linear
from .ops import ACT_LRELU, EPI_MUL_AUXP1


def color_block(in_filters, out_filters, normalization=False):
    """models/LiteISP.py:23-30"""
    layers = [Conv2d(in_filters, out_filters, 1, stride=1, padding=0),
              nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=True), nn.LeakyReLU(0.2)]
    if normalization:
        layers.append(nn.InstanceNorm2d(out_filters, affine=True))
    return layers


class Color_Condition_GFM(nn.Module):
    """Global colour-condition vector (models/LiteISP.py:345-361).  Inference only (Dropout = identity)."""

    def __init__(self, in_channels=4, out_c=32):
        super().__init__()
        self.model = nn.Sequential(
            *color_block(in_channels, 16, normalization=True), *color_block(16, 32, normalization=True),
            *color_block(32, 64, normalization=True), *color_block(64, 128, normalization=True),
            *color_block(128, 128), nn.Dropout(p=0.5), Conv2d(128, out_c, 1, stride=1, padding=0),
            nn.AdaptiveAvgPool2d(1))

    def _f(self, x):
        """NHWC in -> (N,1,1,out_c)"""
        mods = list(self.model)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Conv2d) and i + 2 < len(mods) and isinstance(mods[i + 1], nn.AvgPool2d):
                x = ops.avgpool3s2_lrelu(m._f(x), mods[i + 2].negative_slope)
                i += 3
            elif isinstance(m, nn.InstanceNorm2d):
                x = ops.instance_norm(x, m.weight, m.bias, m.eps)
                i += 1
            elif isinstance(m, nn.Dropout):
                i += 1
            elif isinstance(m, Conv2d):
                x = m._f(x)
                i += 1
            elif isinstance(m, nn.AdaptiveAvgPool2d):
                x = ops.channel_mean(x)
                i += 1
            else:
                raise NotImplementedError(type(m).__name__)
        return x

    def forward(self, img_input):
        v = self._f(ops.to_nhwc(img_input))           # (N,1,1,C)
        return v.permute(0, 3, 1, 2).contiguous()    # (N,C,1,1) like the reference


class Lens_Shading_Correction(nn.Module):
    """Per-pixel coordinate MLP (models/LiteISP.py:363-378)."""

    def __init__(self, in_channels=2, out_c=32, nf=32):
        super().__init__()
        self.model = nn.Sequential(
            Conv2d(in_channels, nf, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            Conv2d(nf, nf, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            Conv2d(nf, nf, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            Conv2d(nf, out_c, 1, 1))

    def _f(self, x):
        sp = None
        for i in (0, 2, 4):   # per-pixel MLP: the hidden maps only exist as the next layer's bf16 planes
            x, sp = self.model[i]._f(x, act=ACT_LRELU, slope=0.1, presplit=sp, emit_split=True, keep_fp32=False)
        return self.model[6]._f(x, presplit=sp)

    def forward(self, img_input):
        return ops.to_nchw(self._f(ops.to_nhwc(img_input)))


class Res_GFM(nn.Module):
    """Residual block with global feature modulation (models/LiteISP.py:537-559).
    forward((x, vec)) -> (fea, vec) like the reference."""

    def __init__(self, in_nc=32, chan=32, cond_c=32, out_nc=32, nf=64):
        super().__init__()
        self.conv0 = Conv2d(in_nc, chan, 3, 1, 1)
        self.conv1 = Conv2d(chan, chan, 3, 1, 1)
        self.GFM_scale_conv0 = Linear(cond_c, nf)
        self.GFM_scale_conv1 = Linear(nf, chan)
        self.GFM_shift_conv0 = Linear(cond_c, nf)
        self.GFM_shift_conv1 = Linear(nf, chan)
        self.out_nc = chan
        self.act = nn.LeakyReLU(inplace=True)

    def _f(self, x, vec):
        """x NHWC, vec (N,1,1,cond_c)"""
        scale = self.GFM_scale_conv1._f(self.GFM_scale_conv0._f(vec, act=ACT_LRELU, slope=0.1))
        shift = self.GFM_shift_conv1._f(self.GFM_shift_conv0._f(vec, act=ACT_LRELU, slope=0.1))
        fea, sp = self.conv0._f(x, cscale=scale.reshape(-1), cshift=shift.reshape(-1), act=ACT_LRELU, slope=0.01,
                                emit_split=True, keep_fp32=False)
        return self.conv1._f(fea, res=x, presplit=sp)

    def forward(self, x):
        fea = self._f(ops.to_nhwc(x[0]), x[1].reshape(x[1].shape[0], 1, 1, -1).contiguous())
        return ops.to_nchw(fea), x[1]


class LiteISPNet_GFM_LSC(nn.Module):
    """models/LiteISP.py:1924-2035."""

    def __init__(self):
        super().__init__()
        ch_1, ch_2, ch_3, n_blocks, cond_c = 48, 128, 128, 4, 32
        self.classifier = Color_Condition_GFM(in_channels=4, out_c=cond_c)
        self.head = N.seq(N.conv(4, ch_1, mode='C'))
        self.lsc = Lens_Shading_Correction(in_channels=2, out_c=ch_1, nf=ch_1)
        self.encoder_modulation1 = N.seq(*[Res_GFM(in_nc=ch_1, chan=ch_1, cond_c=cond_c, out_nc=ch_1, nf=ch_1)])
        self.down1 = N.seq(N.conv(ch_1, ch_1, mode='C'), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                           N.conv(ch_1, ch_1, mode='C'), N.DWTForward(ch_1))
        self.encoder_modulation2 = N.seq(*[Res_GFM(in_nc=ch_1 * 4, chan=ch_1 * 4, cond_c=cond_c, out_nc=ch_1 * 4, nf=ch_1)])
        self.down2 = N.seq(N.conv(ch_1 * 4, ch_1, mode='C'), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                           N.DWTForward(ch_1))
        self.encoder_modulation3 = N.seq(*[Res_GFM(in_nc=ch_1 * 4, chan=ch_1 * 4, cond_c=cond_c, out_nc=ch_1 * 4, nf=ch_1)])
        self.down3 = N.seq(N.conv(ch_1 * 4, ch_2, mode='C'), N.RCAGroup(in_channels=ch_2, out_channels=ch_2, nb=n_blocks),
                           N.DWTForward(ch_2))
        self.encoder_modulation4 = N.seq(*[Res_GFM(in_nc=ch_2 * 4, chan=ch_2 * 4, cond_c=cond_c, out_nc=ch_2 * 4, nf=ch_2)])
        self.middle = N.seq(N.conv(ch_2 * 4, ch_3, mode='C'), N.RCAGroup(in_channels=ch_3, out_channels=ch_3, nb=n_blocks),
                            N.RCAGroup(in_channels=ch_3, out_channels=ch_3, nb=n_blocks), N.conv(ch_3, ch_2 * 4, mode='C'))
        self.up3 = N.seq(N.DWTInverse(ch_2 * 4), N.RCAGroup(in_channels=ch_2, out_channels=ch_2, nb=n_blocks),
                         N.conv(ch_2, ch_1 * 4, mode='C'))
        self.up2 = N.seq(N.DWTInverse(ch_1 * 4), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                         N.conv(ch_1, ch_1 * 4, mode='C'))
        self.up1 = N.seq(N.DWTInverse(ch_1 * 4), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                         N.conv(ch_1, ch_1, mode='C'))
        self.tail = N.seq(N.conv(ch_1, ch_1 * 4, mode='C'), nn.PixelShuffle(upscale_factor=2), N.conv(ch_1, 3, mode='C'))

    def forward(self, x):
        raw, cond, coord = ops.to_nhwc(x[0]), ops.to_nhwc(x[1]), ops.to_nhwc(x[2])
        lsc_fea = self.lsc._f(coord)
        h = self.head._f(raw, epi=EPI_MUL_AUXP1, aux=lsc_fea)              # head(x) * (lsc + 1)
        vec = self.classifier._f(cond)
        h = self.encoder_modulation1._f(h, vec)
        d1 = self.down1._f(h)
        d2 = self.down2._f(self.encoder_modulation2._f(d1, vec))
        d3 = self.down3._f(self.encoder_modulation3._f(d2, vec))
        d4 = self.encoder_modulation4._f(d3, vec)
        m = self.middle._f(d4, res=d3)
        u3 = self.up3._f(m, res=d2)
        u2 = self.up2._f(u3, res=d1)
        u1 = self.up1._f(u2, res=h)
        t = self.tail[0]._f(u1, store=ops.STORE_PS2)
        return self.tail[2]._f(t, store=ops.STORE_NCHW)


class LiteISPNet(nn.Module):
    """models/LiteISP.py:2322-2412: the plain LiteISP UNet -- LiteISPNet_GFM_LSC without colour condition, lens shading and
    modulation blocks (ch_1 = 64).  forward(x) reads x[0], the packed RAW tile (N,4,H,W), and returns (N,3,2H,2W)."""

    def __init__(self):
        super().__init__()
        ch_1, ch_2, ch_3, n_blocks = 64, 128, 128, 4
        self.head = N.seq(N.conv(4, ch_1, mode='C'))
        self.down1 = N.seq(N.conv(ch_1, ch_1, mode='C'), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                           N.conv(ch_1, ch_1, mode='C'), N.DWTForward(ch_1))
        self.down2 = N.seq(N.conv(ch_1 * 4, ch_1, mode='C'), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                           N.DWTForward(ch_1))
        self.down3 = N.seq(N.conv(ch_1 * 4, ch_2, mode='C'), N.RCAGroup(in_channels=ch_2, out_channels=ch_2, nb=n_blocks),
                           N.DWTForward(ch_2))
        self.middle = N.seq(N.conv(ch_2 * 4, ch_3, mode='C'), N.RCAGroup(in_channels=ch_3, out_channels=ch_3, nb=n_blocks),
                            N.RCAGroup(in_channels=ch_3, out_channels=ch_3, nb=n_blocks), N.conv(ch_3, ch_2 * 4, mode='C'))
        self.up3 = N.seq(N.DWTInverse(ch_2 * 4), N.RCAGroup(in_channels=ch_2, out_channels=ch_2, nb=n_blocks),
                         N.conv(ch_2, ch_1 * 4, mode='C'))
        self.up2 = N.seq(N.DWTInverse(ch_1 * 4), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                         N.conv(ch_1, ch_1 * 4, mode='C'))
        self.up1 = N.seq(N.DWTInverse(ch_1 * 4), N.RCAGroup(in_channels=ch_1, out_channels=ch_1, nb=n_blocks),
                         N.conv(ch_1, ch_1, mode='C'))
        self.tail = N.seq(N.conv(ch_1, ch_1 * 4, mode='C'), nn.PixelShuffle(upscale_factor=2), N.conv(ch_1, 3, mode='C'))

    def forward(self, x):
        h = self.head._f(ops.to_nhwc(x[0]))
        d1 = self.down1._f(h)
        d2 = self.down2._f(d1)
        d3 = self.down3._f(d2)
        m = self.middle._f(d3, res=d3)
        u3 = self.up3._f(m, res=d2)
        u2 = self.up2._f(u3, res=d1)
        u1 = self.up1._f(u2, res=h)
        t = self.tail[0]._f(u1, store=ops.STORE_PS2)
        return self.tail[2]._f(t, store=ops.STORE_NCHW)
