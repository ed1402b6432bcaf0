"""Host-side mirror of the reference's ``models/LiteISP.py`` (hot-path classes).

  color_block 23-30, Color_Condition_GFM 345-361, Lens_Shading_Correction 363-378,
  Res_GFM 537-559, LiteISPNet_GFM_LSC 1924-2035 (BASELINE config 1), LiteISPNet 2322-2412,
  ISPUNet_GFM_LSC 1228-1381, ResUNet 2038-2146, MWISP 2149-2218 (the other ISP variants, same kernels re-wired).
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

    def _f(self, x, nchw=False):
        """nchw=True: the result is written once, as the contiguous NCHW map the API returns (`lsc` of forward()); consumers read
        it through conv2d(aux=..., aux_nchw=True)."""
        sp = None
        for i in (0, 2, 4):   # per-pixel MLP: the hidden maps only exist as the next layer's bf16 planes
            x, sp = self.model[i]._f(x, act=ACT_LRELU, slope=0.1, presplit=sp, emit_split=True, keep_fp32=False)
        return self.model[6]._f(x, presplit=sp, store=ops.STORE_NCHW if nchw else ops.STORE_NHWC)

    def layers(self):
        return [self.model[i] for i in (0, 2, 4, 6)]

    def _f_fused(self, coord_nchw, raw=None, conv_first=None, emit_stride=2):
        """One kernel for the whole MLP (+ conv_first(raw) * (lsc + 1) of models/raw2bit.py:1780): hidden maps stay in tensor memory.
        Returns (lsc NCHW, operand planes of the product | None).  Callers check ops.fused_ingest_ok first."""
        return ops.ingest_fused(coord_nchw, self.layers(), 0.1, raw=raw, conv_first=conv_first, emit_stride=emit_stride)

    def forward(self, img_input):
        if ops.fused_ingest_ok(self.layers(), img_input):
            return self._f_fused(img_input)[0]
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

    def _f(self, x, vec, presplit=None, emit_split=False):
        """x NHWC, vec (N,1,1,cond_c); presplit: operand planes of x from its producer; emit_split: returns (out, planes of out)"""
        scale = self.GFM_scale_conv1._f(self.GFM_scale_conv0._f(vec, act=ACT_LRELU, slope=0.1))
        shift = self.GFM_shift_conv1._f(self.GFM_shift_conv0._f(vec, act=ACT_LRELU, slope=0.1))
        fea, sp = self.conv0._f(x, cscale=scale.reshape(-1), cshift=shift.reshape(-1), act=ACT_LRELU, slope=0.01,
                                emit_split=True, keep_fp32=False, presplit=presplit)
        return self.conv1._f(fea, res=x, presplit=sp, emit_split=emit_split)

    def forward(self, x):
        fea = self._f(ops.to_nhwc(x[0]), x[1].reshape(x[1].shape[0], 1, 1, -1).contiguous())
        return ops.to_nchw(fea), x[1]


class LiteISPNet_GFM_LSC(ops.GraphReplay, nn.Module):
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
        """forward([raw, cond, coord]); replayed as one CUDA graph after enable_cuda_graphs() (ops.GraphReplay)"""
        return self._graph_call(self._forward, list(x))

    def _forward(self, x):
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


class LiteISPNet(ops.GraphReplay, nn.Module):
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
        return self._graph_call(self._forward, list(x))

    def _forward(self, x):
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


# ----------------------------------------------------------------------------- the other ISP variants
def _modulate(mods, x, vec):
    """N.seq(*[Res_GFM ...]) called on the (fea, vec) tuple: one block (seq collapses it) or a Sequential of them."""
    for m in ([mods] if isinstance(mods, Res_GFM) else list(mods)):
        x = m._f(x, vec)
    return x


class _UNetISP(ops.GraphReplay, nn.Module):
    """Shared wiring of ISPUNet_GFM_LSC and ResUNet: 3-level UNet of RCAGroups with learned 2x2 stride-2 down-samplers and
    1x1 conv + PixelShuffle up-samplers, chan = 32 -> 64 -> 128 -> 256."""

    def _build(self, chan, n_blocks, modulation):
        cond_c, mb = modulation if modulation else (None, 0)

        def mod(c):
            return N.seq(*[Res_GFM(in_nc=c, chan=c, cond_c=cond_c, out_nc=c, nf=c * 2) for _ in range(mb)])

        self.intro = N.seq(Conv2d(4, chan, 3, 1, 1))
        if modulation:
            self.lsc = Lens_Shading_Correction(in_channels=2, out_c=chan, nf=chan)
            self.encoder_modulation1 = mod(chan)
        self.encoder1 = N.seq(N.RCAGroup(in_channels=chan, out_channels=chan, nb=n_blocks), Conv2d(chan, chan, 3, 1, 1),
                              nn.LeakyReLU(negative_slope=1e-1, inplace=True))
        self.down1 = N.Down2x2(chan, chan * 2)
        chan *= 2
        if modulation:
            self.encoder_modulation2 = mod(chan)
        self.encoder2 = N.seq(N.RCAGroup(in_channels=chan, out_channels=chan, nb=n_blocks), Conv2d(chan, chan, 3, 1, 1),
                              nn.LeakyReLU(negative_slope=1e-1, inplace=True))
        self.down2 = N.Down2x2(chan, chan * 2)
        chan *= 2
        if modulation:
            self.encoder_modulation3 = mod(chan)
        self.encoder3 = N.seq(Conv2d(chan, chan, 3, 1, 1), N.RCAGroup(in_channels=chan, out_channels=chan, nb=n_blocks),
                              Conv2d(chan, chan, 3, 1, 1), nn.LeakyReLU(negative_slope=1e-1, inplace=True))
        self.down3 = N.Down2x2(chan, chan * 2)
        chan *= 2
        if modulation:
            self.middle_modulation = mod(chan)
        self.middle = N.seq(Conv2d(chan, chan, 3, 1, 1), N.RCAGroup(in_channels=chan, out_channels=chan, nb=n_blocks * 2),
                            Conv2d(chan, chan, 3, 1, 1))
        for lvl in (3, 2, 1):
            setattr(self, f"up{lvl}", N.seq(Conv2d(chan, chan * 2, 1, bias=False), nn.PixelShuffle(2)))
            chan //= 2
            if modulation:
                setattr(self, f"decoder_modulation{lvl}", mod(chan))
            setattr(self, f"decoder{lvl}", N.seq(N.RCAGroup(in_channels=chan, out_channels=chan, nb=n_blocks), N.conv(chan, chan, mode='C')))
        self.tail = N.seq(N.conv(chan, chan * 4, mode='C'), nn.PixelShuffle(upscale_factor=2), N.conv(chan, 3, mode='C'))
        self._modulated = bool(modulation)

    def forward(self, x):
        return self._graph_call(self._forward, list(x))

    def _forward(self, x):
        raw = ops.to_nhwc(x[0])
        vec = None
        if self._modulated:
            lsc_fea = self.lsc._f(ops.to_nhwc(x[2]))
            fea_intro = self.intro._f(raw, epi=EPI_MUL_AUXP1, aux=lsc_fea)          # intro(x) * (lsc + 1)
            vec = self.classifier._f(ops.to_nhwc(x[1]))
        else:
            fea_intro = self.intro._f(raw)
        m = (lambda name, t: _modulate(getattr(self, name), t, vec)) if self._modulated else (lambda name, t: t)
        d1 = self.down1._f(self.encoder1._f(m("encoder_modulation1", fea_intro)))
        d2 = self.down2._f(self.encoder2._f(m("encoder_modulation2", d1)))
        d3 = self.down3._f(self.encoder3._f(m("encoder_modulation3", d2)))
        mid = self.middle._f(m("middle_modulation", d3), res=d3)
        u = mid
        for lvl, skip in ((3, d2), (2, d1), (1, fea_intro)):
            u = getattr(self, f"decoder{lvl}")._f(getattr(self, f"up{lvl}")._f(u))
            if self._modulated:          # u = modulation(u) + skip: the block's own residual occupies the conv epilogue
                u = ops.add(m(f"decoder_modulation{lvl}", u), skip)
            else:
                u = ops.add(u, skip)
        t = self.tail[0]._f(u, store=ops.STORE_PS2)
        return self.tail[2]._f(t, store=ops.STORE_NCHW)


class ISPUNet_GFM_LSC(_UNetISP):
    """models/LiteISP.py:1228-1381.  forward([raw, cond, coord]) -> (B,3,2H,2W)."""

    def __init__(self, cond_c=32, chan=32, m_blocks=2):
        super().__init__()
        self.classifier = Color_Condition_GFM(in_channels=4, out_c=cond_c)
        self._build(chan, 2, (cond_c, m_blocks))


class ResUNet(_UNetISP):
    """models/LiteISP.py:2038-2146: ISPUNet without colour condition, lens shading and modulation.  forward(x) reads x[0]."""

    def __init__(self):
        super().__init__()
        self._build(32, 2, None)


class MWISP(ops.GraphReplay, nn.Module):
    """models/LiteISP.py:2149-2218 (multi-level wavelet ISP: Haar DWT / IDWT around RCAGroups of 20 blocks, PReLU activations).
    forward(x, c=None) reads x[0] (N,4,H,W) and returns (N,3,2H,2W)."""

    def __init__(self):
        super().__init__()
        c1, c2, c3, n_b = 64, 128, 128, 20
        self.head = N.DWTForward_()
        self.down1 = N.seq(Conv2d(4 * 4, c1, 3, 1, 1), nn.PReLU(), N.RCAGroup(in_channels=c1, out_channels=c1, nb=n_b))
        self.down2 = N.seq(N.DWTForward_(), Conv2d(c1 * 4, c2, 3, 1, 1), nn.PReLU(), N.RCAGroup(in_channels=c2, out_channels=c2, nb=n_b))
        self.down3 = N.seq(N.DWTForward_(), Conv2d(c2 * 4, c3, 3, 1, 1), nn.PReLU())
        self.middle = N.seq(N.RCAGroup(in_channels=c3, out_channels=c3, nb=n_b), N.RCAGroup(in_channels=c3, out_channels=c3, nb=n_b))
        self.up1 = N.seq(Conv2d(c3, c2 * 4, 3, 1, 1), nn.PReLU(), N.DWTInverse_())
        self.up2 = N.seq(N.RCAGroup(in_channels=c2, out_channels=c2, nb=n_b), Conv2d(c2, c1 * 4, 3, 1, 1), nn.PReLU(), N.DWTInverse_())
        self.up3 = N.seq(N.RCAGroup(in_channels=c1, out_channels=c1, nb=n_b), Conv2d(c1, 16, 3, 1, 1))
        self.tail = N.seq(N.DWTInverse_(), Conv2d(4, 12, 3, 1, 1), nn.PixelShuffle(upscale_factor=2))

    def forward(self, x, c=None):
        return self._graph_call(self._forward, list(x))

    def _forward(self, x):
        c1 = self.head._f(ops.to_nhwc(x[0]))
        c2 = self.down1._f(c1)
        c3 = self.down2._f(c2)
        c4 = self.down3._f(c3)
        m = self.middle._f(c4)
        c5 = ops.add(self.up1._f(m), c3)
        c6 = ops.add(self.up2._f(c5), c2)
        c7 = self.up3._f(c6, res=c1)
        t = self.tail[0]._f(c7)
        return self.tail[1]._f(t, store=ops.STORE_PS2_NCHW)
