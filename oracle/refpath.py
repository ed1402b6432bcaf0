"""Functional torch-CPU fp32 restatement of the reference forward path (TEST ORACLE ONLY).

Every function takes a reference-named ``state_dict`` (``sd``) plus a key prefix and follows
the reference file:line it cites.  It is validated in the authoring container against the
UNMODIFIED reference modules (tests/test_oracle_vs_reference.py, via oracle/ref_import.py)
and against the committed fixtures in tests/golden/ -- and then serves as the checker for
the CUDA product on the GPU box, where /root/reference does not exist.

The entropy-coding arithmetic (CompressAI) is restated in oracle/cai.py / oracle/rans.py;
that part is "parity unpinned" versus upstream (see those headers).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cai as _cai
from . import rans as _rans

_PED = float((2.0 ** -18) ** 2)


# ----------------------------------------------------------------------------- primitives
def conv(sd, p, x, stride=1):
    """nn.Conv2d with padding = k // 2 (every conv on the path uses 'same' padding)."""
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd.get(p + ".bias"), stride=stride, padding=w.shape[-1] // 2)


def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def layernorm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def gdn(sd, p, x, inverse=False):
    """CompressAI GDN (restated in oracle/cai.py:GDN); used at raw2bit.py:1642,1648,1664-1679."""
    ped = torch.tensor(_PED, dtype=torch.float32)
    bb = torch.tensor((1e-6 + _PED) ** 0.5, dtype=torch.float32)
    gb = torch.tensor((0.0 + _PED) ** 0.5, dtype=torch.float32)
    beta = torch.max(sd[p + ".beta"], bb) ** 2 - ped
    gamma = torch.max(sd[p + ".gamma"], gb) ** 2 - ped
    C = x.shape[1]
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


def rb_with_stride(sd, p, x, stride=2):
    """compressai.layers.ResidualBlockWithStride (call sites raw2bit.py:1642,1648,1654,1688)."""
    out = conv(sd, p + ".conv1", x, stride)
    out = F.leaky_relu(out, 0.01)
    out = gdn(sd, p + ".gdn", conv(sd, p + ".conv2", out))
    skip = conv(sd, p + ".skip", x, stride) if (p + ".skip.weight") in sd else x
    return out + skip


def subpel(sd, p, x, r=2):
    """compressai.layers.subpel_conv3x3 = Sequential(conv3x3, PixelShuffle)."""
    return F.pixel_shuffle(conv(sd, p + ".0", x), r)


def rb_upsample(sd, p, x):
    """compressai.layers.ResidualBlockUpsample (raw2bit.py:1664-1670,1686,1701,1710)."""
    out = F.leaky_relu(subpel(sd, p + ".subpel_conv", x), 0.01)
    out = gdn(sd, p + ".igdn", conv(sd, p + ".conv", out), inverse=True)
    return out + subpel(sd, p + ".upsample", x)


def residual_block(sd, p, x):
    """compressai.layers.ResidualBlock (tcm.py:258, raw2bit.py:1681)."""
    out = F.leaky_relu(conv(sd, p + ".conv1", x), 0.01)
    out = F.leaky_relu(conv(sd, p + ".conv2", out), 0.01)
    skip = conv(sd, p + ".skip", x) if (p + ".skip.weight") in sd else x
    return out + skip


# ----------------------------------------------------------------------------- Swin window attention
def _rel_bias(params, ws):
    """tcm.py:209-212 -- bias[h, p, q] = params[h, pi - qi + ws - 1, pj - qj + ws - 1]."""
    ii = torch.arange(ws).repeat_interleave(ws)
    jj = torch.arange(ws).repeat(ws)
    return params[:, ii[:, None] - ii[None, :] + ws - 1, jj[:, None] - jj[None, :] + ws - 1]


def wmsa(sd, p, x, head_dim, ws, shifted):
    """tcm.py:179-207 (WMSA.forward); x is (B, H, W, C)."""
    B, H, W, C = x.shape
    nh = C // head_dim
    sh = ws // 2
    if shifted:
        x = torch.roll(x, shifts=(-sh, -sh), dims=(1, 2))
    nwh, nww = H // ws, W // ws
    xw = x.reshape(B, nwh, ws, nww, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, nwh * nww, ws * ws, C)
    qkv = linear(sd, p + ".embedding_layer", xw)  # (B, nW, P, 3C) ordered [q heads | k heads | v heads]
    qkv = qkv.reshape(B, nwh * nww, ws * ws, 3 * nh, head_dim).permute(3, 0, 1, 2, 4)
    q, k, v = qkv[:nh], qkv[nh:2 * nh], qkv[2 * nh:]
    sim = torch.matmul(q, k.transpose(-1, -2)) * (head_dim ** -0.5)  # (nh, B, nW, P, P)
    sim = sim + _rel_bias(sd[p + ".relative_position_params"], ws)[:, None, None]
    if shifted:  # tcm.py:160-177: tokens that wrapped around may only see their own region
        s = ws - sh
        rr = torch.arange(ws).repeat_interleave(ws) >= s  # local row in wrapped part
        cc = torch.arange(ws).repeat(ws) >= s
        mask = torch.zeros(nwh, nww, ws * ws, ws * ws, dtype=torch.bool)
        mask[-1] |= rr[:, None] != rr[None, :]
        mask[:, -1] |= cc[:, None] != cc[None, :]
        sim = sim.masked_fill(mask.reshape(1, 1, nwh * nww, ws * ws, ws * ws), float("-inf"))
    out = torch.matmul(torch.softmax(sim, dim=-1), v)  # (nh, B, nW, P, hd)
    out = out.permute(1, 2, 3, 0, 4).reshape(B, nwh * nww, ws * ws, C)
    out = linear(sd, p + ".linear", out)
    out = out.reshape(B, nwh, nww, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shifted:
        out = torch.roll(out, shifts=(sh, sh), dims=(1, 2))
    return out


def swin_block(sd, p, x, head_dim, ws, shifted):
    """tcm.py:214-236 (Block.forward); NHWC in/out."""
    x = x + wmsa(sd, p + ".msa", layernorm(sd, p + ".ln1", x), head_dim, ws, shifted)
    h = linear(sd, p + ".mlp.2", F.gelu(linear(sd, p + ".mlp.0", layernorm(sd, p + ".ln2", x))))
    return x + h


def conv_trans_block(sd, p, x, head_dim, ws, shifted, cond=None):
    """tcm.py:260-268 (ConvTransBlock) and raw2bit.py:312-328 (ConvTransBlock_mzj when cond given)."""
    C = x.shape[1] // 2
    cx, tx = torch.split(conv(sd, p + ".conv1_1", x), (C, C), dim=1)
    if cond is None:
        cx = residual_block(sd, p + ".conv_block", cx) + cx
    else:
        ident = cx
        # ResidualBlockWithCA, raw2bit.py:275-289 ; CALayer raw2bit.py:238-253
        out = conv(sd, p + ".conv_block.conv2", F.leaky_relu(conv(sd, p + ".conv_block.conv1", cx), 0.01))
        g = out.mean(dim=(2, 3))
        g = torch.sigmoid(F.linear(F.relu(F.linear(g, sd[p + ".conv_block.ca.fc.0.weight"])),
                                   sd[p + ".conv_block.ca.fc.2.weight"]))
        out = out * g[:, :, None, None] + cx
        # SpatialFeatureTransform, raw2bit.py:878-886
        sp = p + ".spatial_transform"
        scale = conv(sd, sp + ".cond_scale.2", F.relu(conv(sd, sp + ".cond_scale.0", cond)))
        shift = conv(sd, sp + ".cond_shift.2", F.relu(conv(sd, sp + ".cond_shift.0", cond)))
        cx = out * scale + shift + out + ident
    tx = swin_block(sd, p + ".trans_block", tx.permute(0, 2, 3, 1), head_dim, ws, shifted).permute(0, 3, 1, 2)
    return x + conv(sd, p + ".conv1_2", torch.cat((cx, tx), dim=1))


def sw_atten(sd, p, x, head_dim=16, ws=8):
    """tcm.py:270-291 / raw2bit.py:186-207 (SWAtten over compressai AttentionBlock)."""
    x = conv(sd, p + ".in_conv", x)
    if x.size(-1) <= ws or x.size(-2) <= ws:  # tcm.py:301-304 pads and never crops (Appendix A)
        raise ValueError("latent map must be larger than the window (reference pads without cropping)")
    t = x.permute(0, 2, 3, 1)
    t = swin_block(sd, p + ".non_local_block.block_1", t, head_dim, ws, False)
    t = swin_block(sd, p + ".non_local_block.block_2", t, head_dim, ws, True)
    z = t.permute(0, 3, 1, 2)

    def unit(pp, v):
        h = F.relu(conv(sd, pp + ".conv.0", v))
        h = F.relu(conv(sd, pp + ".conv.2", h))
        return F.relu(conv(sd, pp + ".conv.4", h) + v)

    a, b = x, z
    for i in range(3):
        a = unit(f"{p}.conv_a.{i}", a)
        b = unit(f"{p}.conv_b.{i}", b)
    b = conv(sd, p + ".conv_b.3", b)
    return conv(sd, p + ".out_conv", a * torch.sigmoid(b) + x)


# ----------------------------------------------------------------------------- conditioning nets
def color_condition_gfm(sd, p, x):
    """LiteISP.py:23-30,345-361 -- 5 x [1x1, AvgPool(3,2,1), LReLU .2, (InstanceNorm)] -> 1x1 -> GAP."""
    idx = 0
    for blk in range(5):
        x = conv(sd, f"{p}.model.{idx}", x)
        x = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=True)
        x = F.leaky_relu(x, 0.2)
        idx += 3
        if blk < 4:
            x = F.instance_norm(x, weight=sd[f"{p}.model.{idx}.weight"], bias=sd[f"{p}.model.{idx}.bias"], eps=1e-5)
            idx += 1
    idx += 1  # Dropout (eval = identity)
    x = conv(sd, f"{p}.model.{idx}", x)
    return x.mean(dim=(2, 3))


def lens_shading(sd, p, coord):
    """LiteISP.py:363-378 -- per-pixel MLP of 1x1 convs, LeakyReLU(0.1)."""
    x = coord
    for i in (0, 2, 4):
        x = F.leaky_relu(conv(sd, f"{p}.model.{i}", x), 0.1)
    return conv(sd, f"{p}.model.6", x)


def res_gfm(sd, p, x, vec):
    """LiteISP.py:537-559."""
    fea = conv(sd, p + ".conv0", x)
    scale = linear(sd, p + ".GFM_scale_conv1", F.leaky_relu(linear(sd, p + ".GFM_scale_conv0", vec), 0.1))
    shift = linear(sd, p + ".GFM_shift_conv1", F.leaky_relu(linear(sd, p + ".GFM_shift_conv0", vec), 0.1))
    fea = fea * scale[:, :, None, None] + shift[:, :, None, None] + fea
    return conv(sd, p + ".conv1", F.leaky_relu(fea, 0.01)) + x


def hybrid_condition(sd, p, x):
    """raw2bit.py:817-858 (HybridConditionModule) with its Enc/Dec/Conv blocks raw2bit.py:730-814."""
    def cb(pp, v, stride=1):
        return F.relu(conv(sd, pp + ".conv", v, stride))

    def enc(pp, v):
        return cb(pp + ".conv", cb(pp + ".down", v, 2))

    def dec(pp, v1, v2):
        v1 = F.interpolate(v1, scale_factor=2, mode="bilinear", align_corners=True)
        v1 = cb(pp + ".up.1", v1)
        return cb(pp + ".conv", torch.cat([v2, v1], dim=1))

    x1 = cb(p + ".in_conv", x)
    x2 = enc(p + ".enc_1", x1)
    x3 = enc(p + ".enc_2", x2)
    x4 = enc(p + ".enc_3", x3)
    y = dec(p + ".dec_1", x4, x3)
    y = dec(p + ".dec_2", y, x2)
    y = dec(p + ".dec_3", y, x1)
    y = cb(p + ".out_conv", y)
    c1 = conv(sd, p + ".CondNet1.2", F.leaky_relu(conv(sd, p + ".CondNet1.0", y, 2), 0.1))
    c2 = conv(sd, p + ".CondNet2.2", F.leaky_relu(conv(sd, p + ".CondNet2.0", y, 2), 0.1), 2)
    c3 = F.leaky_relu(conv(sd, p + ".CondNet3.0", y, 2), 0.1)
    c3 = F.leaky_relu(conv(sd, p + ".CondNet3.2", c3, 2), 0.1)
    c3 = conv(sd, p + ".CondNet3.4", c3, 2)
    return [c1, c2, c3]


# ----------------------------------------------------------------------------- entropy models
def _eb(sd, p="entropy_bottleneck"):
    """Instantiates the restated EntropyBottleneck with the weights of ``sd``."""
    C = sd[p + ".quantiles"].shape[0]
    eb = _cai.EntropyBottleneck(C)
    own = eb.state_dict()
    for k in own:
        if (p + "." + k) in sd and k not in ("_offset", "_quantized_cdf", "_cdf_length"):
            own[k].copy_(sd[p + "." + k])
    eb.eval()
    return eb


_GC_CACHE = []


def _gc():
    """GaussianConditional with the default scale table; the tables are constants, so update() runs once."""
    if not _GC_CACHE:
        gc = _cai.GaussianConditional(None)
        gc.update_scale_table(_cai.get_scale_table())
        gc.eval()
        _GC_CACHE.append(gc)
    return _GC_CACHE[0]


# ----------------------------------------------------------------------------- raw_compression_tcm_final
HEAD_DIM = (8, 16, 32, 32, 16, 8)


def analysis(sd, x):
    """raw2bit.py:1771-1796 (identical lines in compress(): 1877-1901). Returns y and by-products."""
    raw, cond_img, coord = x
    fea = conv(sd, "conv_first", raw)
    vec = color_condition_gfm(sd, "classifier", cond_img)
    lsc = lens_shading(sd, "lsc", coord)
    local = hybrid_condition(sd, "local_condition", raw)
    fea = fea * (lsc + 1)
    fea = rb_with_stride(sd, "conv_down", fea)
    for lvl in range(3):
        fea = res_gfm(sd, f"gfm{lvl + 1}.0", fea, vec)
        for i in range(2):
            fea = conv_trans_block(sd, f"m_down{lvl + 1}.{i}", fea, HEAD_DIM[lvl], 8, i % 2 == 1, cond=local[lvl])
        if lvl < 2:
            fea = rb_with_stride(sd, f"m_down{lvl + 1}_down", fea)
    y = conv(sd, "m_down3_down", fea, 2)
    return y, lsc, local, vec


def hyper_analysis(sd, y):
    """h_a, raw2bit.py:1688-1695."""
    z = rb_with_stride(sd, "h_a.0", y)
    for i in range(2):
        z = conv_trans_block(sd, f"h_a.{i + 1}", z, 32, 4, i % 2 == 1)
    return conv(sd, "h_a.3", z, 2)


def hyper_synthesis(sd, p, z_hat):
    """h_mean_s / h_scale_s, raw2bit.py:1697-1715."""
    h = rb_upsample(sd, p + ".0", z_hat)
    for i in range(2):
        h = conv_trans_block(sd, f"{p}.{i + 1}", h, 32, 4, i % 2 == 1)
    return subpel(sd, p + ".3", h)


def synthesis(sd, y_hat):
    """g_s, raw2bit.py:1664-1686."""
    h = rb_upsample(sd, "g_s.0", y_hat)
    idx = 1
    for lvl in range(3):
        for i in range(2):
            h = conv_trans_block(sd, f"g_s.{idx}", h, HEAD_DIM[3 + lvl], 8, i % 2 == 1)
            idx += 1
        h = rb_upsample(sd, f"g_s.{idx}", h) if lvl < 2 else subpel(sd, f"g_s.{idx}", h)
        idx += 1
    h = residual_block(sd, f"g_s.{idx}", h)
    return subpel(sd, f"g_s.{idx + 1}", h)


def _cc(sd, p, x):
    """cc_mean/cc_scale/lrp transform: 3x3 -> GELU -> 3x3 -> GELU -> 3x3 (raw2bit.py:1727-1754)."""
    return conv(sd, p + ".4", F.gelu(conv(sd, p + ".2", F.gelu(conv(sd, p + ".0", x)))))


def slice_params(sd, i, latent_means, latent_scales, y_hat_slices):
    """Per-slice entropy parameters, raw2bit.py:1818-1828."""
    mean_support = sw_atten(sd, f"atten_mean.{i}.0", torch.cat([latent_means] + y_hat_slices, dim=1))
    mu = _cc(sd, f"cc_mean_transforms.{i}", mean_support)
    scale_support = sw_atten(sd, f"atten_scale.{i}.0", torch.cat([latent_scales] + y_hat_slices, dim=1))
    scale = _cc(sd, f"cc_scale_transforms.{i}", scale_support)
    return mean_support, mu, scale


def lrp(sd, i, mean_support, y_hat_slice):
    """raw2bit.py:1835-1838."""
    return 0.5 * torch.tanh(_cc(sd, f"lrp_transforms.{i}", torch.cat([mean_support, y_hat_slice], dim=1)))


def _forward_core(sd, y, num_slices=5, trace=None):
    """Hyper-prior + slice loop shared by TCM.forward (tcm.py:437-481) and raw_compression_tcm_final.forward
    (raw2bit.py:1798-1846); returns (y_hat, y_likelihoods, z_likelihoods, means, scales)."""
    z = hyper_analysis(sd, y)
    eb, gc = _eb(sd), _gc()
    _, z_lik = eb(z)
    med = eb._get_medians().reshape(1, -1, 1, 1)
    z_hat = torch.round(z - med) + med
    latent_scales = hyper_synthesis(sd, "h_scale_s", z_hat)
    latent_means = hyper_synthesis(sd, "h_mean_s", z_hat)
    y_hat_slices, liks, mus, scales = [], [], [], []
    for i, y_slice in enumerate(y.chunk(num_slices, 1)):
        mean_support, mu, scale = slice_params(sd, i, latent_means, latent_scales, y_hat_slices)
        _, lik = gc(y_slice, scale, mu)
        y_hat = torch.round(y_slice - mu) + mu
        y_hat = y_hat + lrp(sd, i, mean_support, y_hat)
        y_hat_slices.append(y_hat)
        liks.append(lik), mus.append(mu), scales.append(scale)
    y_hat = torch.cat(y_hat_slices, dim=1)
    if trace is not None:
        trace.update(z=z, z_hat=z_hat, latent_means=latent_means, latent_scales=latent_scales, y_hat=y_hat)
    return y_hat, torch.cat(liks, dim=1), z_lik, torch.cat(mus, dim=1), torch.cat(scales, dim=1)


def _compress_core(sd, y, num_slices=5, trace=None):
    """tcm.py:515-570 / raw2bit.py:1903-1960."""
    z = hyper_analysis(sd, y)
    eb, gc = _eb(sd), _gc()
    eb.update(force=True)
    z_strings = eb.compress(z)
    z_hat = eb.decompress(z_strings, z.size()[-2:])
    latent_scales = hyper_synthesis(sd, "h_scale_s", z_hat)
    latent_means = hyper_synthesis(sd, "h_mean_s", z_hat)
    symbols, indexes, y_hat_slices = [], [], []
    for i, y_slice in enumerate(y.chunk(num_slices, 1)):
        mean_support, mu, scale = slice_params(sd, i, latent_means, latent_scales, y_hat_slices)
        index = gc.build_indexes(scale)
        y_q = gc.quantize(y_slice, "symbols", mu)
        y_hat = y_q + mu
        symbols.append(y_q.reshape(-1))
        indexes.append(index.reshape(-1))
        y_hat = y_hat + lrp(sd, i, mean_support, y_hat)
        y_hat_slices.append(y_hat)
    symbols = torch.cat(symbols).numpy().astype(np.int32)
    indexes = torch.cat(indexes).numpy().astype(np.int32)
    y_string = encode_stream(symbols, indexes, gc)
    if trace is not None:
        trace.update(symbols=symbols, indexes=indexes, z=z, y=y)
    return {"strings": [[y_string], z_strings], "shape": z.size()[-2:]}


def _decompress_core(sd, strings, shape, num_slices=5):
    """tcm.py:592-634 / raw2bit.py:1982-2024; returns y_hat."""
    eb, gc = _eb(sd), _gc()
    eb.update(force=True)
    z_hat = eb.decompress(strings[1], shape)
    latent_scales = hyper_synthesis(sd, "h_scale_s", z_hat)
    latent_means = hyper_synthesis(sd, "h_mean_s", z_hat)
    h, w = z_hat.shape[2] * 4, z_hat.shape[3] * 4
    dec = StreamDecoder(strings[0][0], gc)
    y_hat_slices = []
    for i in range(num_slices):
        mean_support, mu, scale = slice_params(sd, i, latent_means, latent_scales, y_hat_slices)
        index = gc.build_indexes(scale)
        rv = dec.decode(index.reshape(-1).numpy().astype(np.int32))
        rv = torch.from_numpy(rv.astype(np.float32)).reshape(1, -1, h, w)
        y_hat = rv + mu
        y_hat = y_hat + lrp(sd, i, mean_support, y_hat)
        y_hat_slices.append(y_hat)
    return torch.cat(y_hat_slices, dim=1)


@torch.no_grad()
def final_forward(sd, x, num_slices=5, trace=None):
    """raw_compression_tcm_final.forward, raw2bit.py:1766-1855 (eval mode)."""
    y, lsc, local, vec = analysis(sd, x)
    y_hat, y_lik, z_lik, means, scales = _forward_core(sd, y, num_slices, trace)
    x_hat = synthesis(sd, y_hat)
    if trace is not None:
        trace.update(gfm_vector=vec, local=local)
    return {"x_hat": x_hat, "y": y, "lft": local[2], "lsc": lsc, "likelihoods": {"y": y_lik, "z": z_lik},
            "para": {"means": means, "scales": scales, "y": y}}


@torch.no_grad()
def final_compress(sd, x, num_slices=5, trace=None):
    """raw_compression_tcm_final.compress, raw2bit.py:1876-1960."""
    return _compress_core(sd, analysis(sd, x)[0], num_slices, trace)


@torch.no_grad()
def final_decompress(sd, strings, shape, num_slices=5):
    """raw_compression_tcm_final.decompress, raw2bit.py:1982-2027."""
    return {"x_hat": synthesis(sd, _decompress_core(sd, strings, shape, num_slices)).clamp_(0, 1)}


# ----------------------------------------------------------------------------- TCM (RGB baseline, tcm.py:320-637)
def tcm_analysis(sd, x):
    """g_a, tcm.py:335-347,360: RBWS(3,128) + 3 x (2 ConvTransBlock + down)."""
    h = rb_with_stride(sd, "g_a.0", x)
    idx = 1
    for lvl in range(3):
        for i in range(2):
            h = conv_trans_block(sd, f"g_a.{idx}", h, HEAD_DIM[lvl], 8, i % 2 == 1)
            idx += 1
        h = rb_with_stride(sd, f"g_a.{idx}", h) if lvl < 2 else conv(sd, f"g_a.{idx}", h, 2)
        idx += 1
    return h


def tcm_synthesis(sd, y_hat):
    """g_s, tcm.py:349-363: RBU(320,128) + 3 x (2 ConvTransBlock + up), last up = subpel_conv3x3(128, 3, 2)."""
    h = rb_upsample(sd, "g_s.0", y_hat)
    idx = 1
    for lvl in range(3):
        for i in range(2):
            h = conv_trans_block(sd, f"g_s.{idx}", h, HEAD_DIM[3 + lvl], 8, i % 2 == 1)
            idx += 1
        h = rb_upsample(sd, f"g_s.{idx}", h) if lvl < 2 else subpel(sd, f"g_s.{idx}", h)
        idx += 1
    return h


@torch.no_grad()
def tcm_forward(sd, x, num_slices=5):
    """TCM.forward, tcm.py:437-490."""
    y = tcm_analysis(sd, x)
    y_hat, y_lik, z_lik, means, scales = _forward_core(sd, y, num_slices)
    return {"x_hat": tcm_synthesis(sd, y_hat), "likelihoods": {"y": y_lik, "z": z_lik},
            "para": {"means": means, "scales": scales, "y": y}}


@torch.no_grad()
def tcm_compress(sd, x, num_slices=5, trace=None):
    """TCM.compress, tcm.py:511-570."""
    return _compress_core(sd, tcm_analysis(sd, x), num_slices, trace)


@torch.no_grad()
def tcm_decompress(sd, strings, shape, num_slices=5):
    """TCM.decompress, tcm.py:592-637."""
    return {"x_hat": tcm_synthesis(sd, _decompress_core(sd, strings, shape, num_slices)).clamp_(0, 1)}


# ----------------------------------------------------------------------------- fast coder (plain C) wrappers
_clib = None


def _c():
    """oracle/rans_c.c compiled on demand into oracle/_build (falls back to pure Python)."""
    global _clib
    if _clib is not None:
        return _clib or None
    import ctypes
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "_build", "librans_oracle.so")
    src = os.path.join(here, "rans_c.c")
    try:
        if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.orc_rans_encode.restype = ctypes.c_longlong
        _clib = lib
    except Exception:
        _clib = False
    return _clib or None


def _tables(gc):
    cdf = np.ascontiguousarray(gc.quantized_cdf.numpy().astype(np.int32))
    sizes = np.ascontiguousarray(gc.cdf_length.numpy().astype(np.int32).reshape(-1))
    offs = np.ascontiguousarray(gc.offset.numpy().astype(np.int32).reshape(-1))
    return cdf, sizes, offs


def encode_stream(symbols, indexes, gc) -> bytes:
    cdf, sizes, offs = _tables(gc)
    lib = _c()
    if lib is None:
        return _rans.encode_with_indexes(symbols, indexes, cdf.tolist(), sizes, offs)
    import ctypes

    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    indexes = np.ascontiguousarray(indexes, dtype=np.int32)
    cap = 8 * len(symbols) + 64
    out = np.empty(cap, dtype=np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n = lib.orc_rans_encode(P(symbols), P(indexes), ctypes.c_longlong(len(symbols)), P(cdf),
                            ctypes.c_int(cdf.shape[1]), P(sizes), P(offs), P(out), ctypes.c_longlong(cap))
    assert n > 0
    return out[:n].tobytes()


class StreamDecoder:
    def __init__(self, stream: bytes, gc):
        self.tab = _tables(gc)
        self.stream = np.frombuffer(stream, dtype=np.uint8).copy()
        self.st = np.array([-1, 0, 0], dtype=np.int64)
        self.py = None if _c() is not None else _rans.Decoder(stream)

    def decode(self, indexes):
        cdf, sizes, offs = self.tab
        if self.py is not None:
            return np.asarray(self.py.decode_stream(indexes, cdf.tolist(), sizes, offs), dtype=np.int32)
        import ctypes

        lib = _c()
        indexes = np.ascontiguousarray(indexes, dtype=np.int32)
        out = np.empty(len(indexes), dtype=np.int32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.orc_rans_decode(P(self.stream), ctypes.c_longlong(len(self.stream)), P(indexes),
                            ctypes.c_longlong(len(indexes)), P(cdf), ctypes.c_int(cdf.shape[1]), P(sizes),
                            P(offs), P(out), P(self.st))
        return out


# ----------------------------------------------------------------------------- LiteISPNet_GFM_LSC
def _rcab(sd, p, x):
    """networks.py:296-311 (RCABlock) with CALayer networks.py:255-270."""
    res = conv(sd, p + ".res.2", F.relu(conv(sd, p + ".res.0", x)))
    g = res.mean(dim=(2, 3), keepdim=True)
    g = torch.sigmoid(conv(sd, p + ".ca.conv_du.2", F.relu(conv(sd, p + ".ca.conv_du.0", g))))
    return res * g + x


def rca_group(sd, p, x, nb=4):
    """networks.py:317-335."""
    h = x
    for i in range(nb):
        h = _rcab(sd, f"{p}.rg.{i}", h)
    return conv(sd, f"{p}.rg.{nb}", h) + x


_HAAR = torch.tensor([[[[0.5, 0.5], [0.5, 0.5]]], [[[0.5, 0.5], [-0.5, -0.5]]],
                      [[[0.5, -0.5], [0.5, -0.5]]], [[[0.5, -0.5], [-0.5, 0.5]]]])


def dwt_forward(x):
    """networks.py:224-235: grouped 2x2 stride-2 conv, output channel order [LL,LH,HL,HH] per input ch."""
    C = x.shape[1]
    return F.conv2d(x, _HAAR.repeat(C, 1, 1, 1), stride=2, groups=C)


def dwt_inverse(x):
    """networks.py:238-249."""
    C = x.shape[1] // 4
    return F.conv_transpose2d(x, _HAAR.repeat(C, 1, 1, 1), stride=2, groups=C)


@torch.no_grad()
def liteisp_gfm_lsc_forward(sd, x):
    """LiteISPNet_GFM_LSC.forward, LiteISP.py:2002-2035."""
    raw, cond_img, coord = x
    h = conv(sd, "head.0", raw) if "head.0.weight" in sd else conv(sd, "head", raw)
    h = h * (lens_shading(sd, "lsc", coord) + 1)
    vec = color_condition_gfm(sd, "classifier", cond_img)
    h = res_gfm(sd, "encoder_modulation1", h, vec)
    d1 = conv(sd, "down1.2", rca_group(sd, "down1.1", conv(sd, "down1.0", h)))
    d1 = dwt_forward(d1)
    d2 = res_gfm(sd, "encoder_modulation2", d1, vec)
    d2 = dwt_forward(rca_group(sd, "down2.1", conv(sd, "down2.0", d2)))
    d3 = res_gfm(sd, "encoder_modulation3", d2, vec)
    d3 = dwt_forward(rca_group(sd, "down3.1", conv(sd, "down3.0", d3)))
    d4 = res_gfm(sd, "encoder_modulation4", d3, vec)
    m = conv(sd, "middle.3", rca_group(sd, "middle.2", rca_group(sd, "middle.1", conv(sd, "middle.0", d4)))) + d3
    u3 = conv(sd, "up3.2", rca_group(sd, "up3.1", dwt_inverse(m))) + d2
    u2 = conv(sd, "up2.2", rca_group(sd, "up2.1", dwt_inverse(u3))) + d1
    u1 = conv(sd, "up1.2", rca_group(sd, "up1.1", dwt_inverse(u2))) + h
    return conv(sd, "tail.2", F.pixel_shuffle(conv(sd, "tail.0", u1), 2))


@torch.no_grad()
def liteisp_forward(sd, x):
    """LiteISPNet.forward, LiteISP.py:2385-2412: the GFM_LSC network without colour condition, lens shading and modulation
    (only x[0], the packed RAW tile, is read)."""
    raw = x[0]
    h = conv(sd, "head.0", raw) if "head.0.weight" in sd else conv(sd, "head", raw)
    d1 = dwt_forward(conv(sd, "down1.2", rca_group(sd, "down1.1", conv(sd, "down1.0", h))))
    d2 = dwt_forward(rca_group(sd, "down2.1", conv(sd, "down2.0", d1)))
    d3 = dwt_forward(rca_group(sd, "down3.1", conv(sd, "down3.0", d2)))
    m = conv(sd, "middle.3", rca_group(sd, "middle.2", rca_group(sd, "middle.1", conv(sd, "middle.0", d3)))) + d3
    u3 = conv(sd, "up3.2", rca_group(sd, "up3.1", dwt_inverse(m))) + d2
    u2 = conv(sd, "up2.2", rca_group(sd, "up2.1", dwt_inverse(u3))) + d1
    u1 = conv(sd, "up1.2", rca_group(sd, "up1.1", dwt_inverse(u2))) + h
    return conv(sd, "tail.2", F.pixel_shuffle(conv(sd, "tail.0", u1), 2))


# ----------------------------------------------------------------------------- the other ISP variants
def _down2x2(sd, p, x):
    """nn.Conv2d(C, 2C, 2, 2): kernel 2, stride 2, no padding (LiteISP.py:1253,1265,1278 / 2056,2065,2075)."""
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=2)


def _unet_isp(sd, x, m_blocks):
    """Shared body of ISPUNet_GFM_LSC.forward (LiteISP.py:1340-1381, m_blocks Res_GFM blocks per modulation stage) and
    ResUNet.forward (LiteISP.py:2122-2146, m_blocks = 0: no condition, no lens shading)."""
    raw = x[0]
    fea_intro = conv(sd, "intro", raw)
    vec = None
    if m_blocks:
        vec = color_condition_gfm(sd, "classifier", x[1])
        fea_intro = fea_intro * (lens_shading(sd, "lsc", x[2]) + 1)

    def mod(p, t):
        if m_blocks == 1:
            return res_gfm(sd, p, t, vec)
        for i in range(m_blocks):
            t = res_gfm(sd, f"{p}.{i}", t, vec)
        return t

    def enc(p, t, lead):
        o = 0
        if lead:
            t = conv(sd, f"{p}.0", t)
            o = 1
        return F.leaky_relu(conv(sd, f"{p}.{o + 1}", rca_group(sd, f"{p}.{o}", t, nb=2)), 0.1)

    d1 = _down2x2(sd, "down1", enc("encoder1", mod("encoder_modulation1", fea_intro), False))
    d2 = _down2x2(sd, "down2", enc("encoder2", mod("encoder_modulation2", d1), False))
    d3 = _down2x2(sd, "down3", enc("encoder3", mod("encoder_modulation3", d2), True))
    mid = mod("middle_modulation", d3)
    mid = conv(sd, "middle.2", rca_group(sd, "middle.1", conv(sd, "middle.0", mid), nb=4)) + d3
    u = mid
    for lvl, skip in ((3, d2), (2, d1), (1, fea_intro)):
        u = F.pixel_shuffle(conv(sd, f"up{lvl}.0", u), 2)
        u = conv(sd, f"decoder{lvl}.1", rca_group(sd, f"decoder{lvl}.0", u, nb=2))
        u = mod(f"decoder_modulation{lvl}", u) + skip
    return conv(sd, "tail.2", F.pixel_shuffle(conv(sd, "tail.0", u), 2))


@torch.no_grad()
def ispunet_gfm_lsc_forward(sd, x, m_blocks=2):
    """ISPUNet_GFM_LSC.forward, LiteISP.py:1228-1381."""
    return _unet_isp(sd, x, m_blocks)


@torch.no_grad()
def resunet_forward(sd, x):
    """ResUNet.forward, LiteISP.py:2038-2146 (reads x[0] only)."""
    return _unet_isp(sd, x, 0)


@torch.no_grad()
def mwisp_forward(sd, x):
    """MWISP.forward, LiteISP.py:2149-2218: DWTForward_/DWTInverse_ (networks.py:10-48, the same Haar bank as DWTForward),
    nn.PReLU() with one shared slope, RCAGroups of 20 blocks."""
    def prelu(p, t):
        return F.prelu(t, sd[p + ".weight"])

    c1 = dwt_forward(x[0])
    c2 = rca_group(sd, "down1.2", prelu("down1.1", conv(sd, "down1.0", c1)), nb=20)
    c3 = rca_group(sd, "down2.3", prelu("down2.2", conv(sd, "down2.1", dwt_forward(c2))), nb=20)
    c4 = prelu("down3.2", conv(sd, "down3.1", dwt_forward(c3)))
    m = rca_group(sd, "middle.1", rca_group(sd, "middle.0", c4, nb=20), nb=20)
    c5 = dwt_inverse(prelu("up1.1", conv(sd, "up1.0", m))) + c3
    c6 = dwt_inverse(prelu("up2.2", conv(sd, "up2.1", rca_group(sd, "up2.0", c5, nb=20)))) + c2
    c7 = conv(sd, "up3.1", rca_group(sd, "up3.0", c6, nb=20)) + c1
    return F.pixel_shuffle(conv(sd, "tail.1", dwt_inverse(c7)), 2)


# ----------------------------------------------------------------------------- GroupMix
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _sepconv(sd, p, x):
    """groupmix.py:240-249 (depthwise k x k then pointwise, both bias-free)."""
    w = sd[p + ".conv1.weight"]
    x = F.conv2d(x, w, None, padding=w.shape[-1] // 2, groups=x.shape[1])
    return F.conv2d(x, sd[p + ".pointwise_conv.weight"])


def efficient_att(sd, p, x, size, num_heads):
    """groupmix.py:177-200 with Aggregator 82-105, Agg_0 47-53, ConvRelPosEnc 138-156."""
    B, N, C = x.shape
    H, W = size
    assert N == H * W
    qkv = linear(sd, p + ".qkv", x).reshape(B, N, 3, C).permute(2, 0, 3, 1).reshape(3, B, C, H, W)
    sdim = C // 5
    ap = p + ".aggregator"
    # segment 4 of q,k,v -> local branch
    loc = qkv[:, :, 4 * sdim:].permute(1, 0, 2, 3, 4).reshape(B, 3 * sdim, H, W)
    loc = _sepconv(sd, ap + ".agg0.conv", loc).reshape(B, sdim, N).permute(0, 2, 1)
    loc = F.hardswish(F.layer_norm(loc, (sdim,), sd[ap + ".agg0.norm.weight"], sd[ap + ".agg0.norm.bias"], 1e-5))
    t = qkv.reshape(3 * B, C, H, W)
    segs = [F.hardswish(_bn(sd, ap + ".norm0", t[:, :sdim]))]
    for j in (1, 2, 3):
        segs.append(F.hardswish(_bn(sd, f"{ap}.norm{j}", _sepconv(sd, f"{ap}.agg{j}", t[:, j * sdim:(j + 1) * sdim]))))
    t = torch.cat(segs, dim=1)  # (3B, 4C/5, H, W)
    Ct = 4 * sdim
    Ch = Ct // num_heads
    t = t.reshape(3, B, num_heads, Ch, N).permute(0, 1, 2, 4, 3)
    q, k, v = t[0], t[1], t[2]  # (B, h, N, Ch)
    ks = k.softmax(dim=2)
    kv = torch.einsum("bhnk,bhnv->bhkv", ks, v)
    eff = torch.einsum("bhnk,bhkv->bhnv", q, kv)
    # crpe: depthwise 3/5/7 conv over v, head groups 2/3/3 (groupmix.py:175)
    vimg = v.permute(0, 1, 3, 2).reshape(B, Ct, H, W)
    outs, c0 = [], 0
    for j, heads in enumerate((2, 3, 3)):
        cj = heads * Ch
        w = sd[f"{p}.crpe.conv_list.{j}.weight"]
        outs.append(F.conv2d(vimg[:, c0:c0 + cj], w, sd[f"{p}.crpe.conv_list.{j}.bias"],
                             padding=w.shape[-1] // 2, groups=cj))
        c0 += cj
    crpe = q * torch.cat(outs, dim=1).reshape(B, num_heads, Ch, N).permute(0, 1, 3, 2)
    scale = (C // num_heads) ** -0.5
    out = (scale * eff + crpe).transpose(1, 2).reshape(B, N, Ct)
    return linear(sd, p + ".proj", torch.cat([out, loc], dim=-1))


@torch.no_grad()
def gma_block(sd, x, size, num_heads=8, p=""):
    """GMA_Block.forward, groupmix.py:289-299 (copy at raw2bit.py:132-142)."""
    pre = (p + ".") if p else ""
    B, N, C = x.shape
    H, W = size
    feat = x.transpose(1, 2).reshape(B, C, H, W)
    w = sd[pre + "cpe.proj.weight"]
    x = (F.conv2d(feat, w, sd[pre + "cpe.proj.bias"], padding=1, groups=C) + feat).flatten(2).transpose(1, 2)
    cur = layernorm(sd, pre + "norm1", x)
    x = x + efficient_att(sd, pre + "att", cur, size, num_heads)
    cur = layernorm(sd, pre + "norm2", x)
    return x + linear(sd, pre + "mlp.fc2", F.gelu(linear(sd, pre + "mlp.fc1", cur)))


def _tokens(x):
    B, C, H, W = x.shape
    return x.flatten(2).transpose(1, 2), (H, W)


def _untokens(t, size):
    B, N, C = t.shape
    return t.transpose(1, 2).reshape(B, C, size[0], size[1])


@torch.no_grad()
def gma_pair(sd, p, x, num_heads):
    """GMABlock.forward, raw2bit.py:168-184 (NCHW in/out, two GMA_Blocks)."""
    pre = (p + ".") if p else ""
    t, size = _tokens(x)
    t = gma_block(sd, t, size, num_heads, pre + "block_1")
    t = gma_block(sd, t, size, num_heads, pre + "block_2")
    return _untokens(t, size)


@torch.no_grad()
def gma_atten(sd, p, x, num_heads):
    """GMAAtten.forward, raw2bit.py:225-234 (compressai AttentionBlock gate around a GMABlock)."""
    pre = (p + ".") if p else ""
    x = conv(sd, pre + "in_conv", x)
    z = gma_pair(sd, pre + "non_local_block", x, num_heads)

    def unit(pp, v):
        h = F.relu(conv(sd, pp + ".conv.0", v))
        h = F.relu(conv(sd, pp + ".conv.2", h))
        return F.relu(conv(sd, pp + ".conv.4", h) + v)

    a, b = x, z
    for i in range(3):
        a = unit(f"{pre}conv_a.{i}", a)
        b = unit(f"{pre}conv_b.{i}", b)
    b = conv(sd, pre + "conv_b.3", b)
    return conv(sd, pre + "out_conv", a * torch.sigmoid(b) + x)


@torch.no_grad()
def conv_gma_block(sd, p, x, conv_dim, num_heads):
    """ConvGMABlock.forward, raw2bit.py:346-355."""
    pre = (p + ".") if p else ""
    both = conv(sd, pre + "conv1_1", x)
    cx, tx = both[:, :conv_dim], both[:, conv_dim:]
    cx = residual_block(sd, pre + "conv_block", cx) + cx
    t, size = _tokens(tx)
    tx = _untokens(gma_block(sd, t, size, num_heads, pre + "trans_block"), size)
    return x + conv(sd, pre + "conv1_2", torch.cat((cx, tx), dim=1))


def psnr(a, b, peak=None):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if peak is None:
        peak = float(b.abs().max())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
