"""Times the packed-Bayer ingest at T (default 2048): the fused kernel (rcn_ingest_fused) against the layer-by-layer path.
usage: prof_ingest.py [T] [fused|layers|both]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from realcamnet_b200 import LiteISP, ops, synthetic
from realcamnet_b200.layers import conv3x3

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
which = sys.argv[2] if len(sys.argv) > 2 else "both"
dev = torch.device("cuda:0")
ops.set_engine("bf16x3")
l = LiteISP.Lens_Shading_Correction(2, 128, 128)
cf = conv3x3(4, 128)
synthetic.fill_(l, seed=6)
synthetic.fill_(cf, seed=7)
l, cf = l.to(dev), cf.to(dev)
g = torch.Generator().manual_seed(0)
coord = (torch.rand(1, 2, T, T, generator=g) * 2 - 1).to(dev)
raw = ops.to_nhwc(torch.rand(1, 4, T, T, generator=g).to(dev))


def fused():
    return l._f_fused(coord, raw, cf, emit_stride=2)


def fused_lsc():
    return l._f_fused(coord)


def layers():
    lw = l._f(ops.to_nhwc(coord), nchw=True)
    return cf._f(raw, epi=ops.EPI_MUL_AUXP1, aux=lw, aux_nchw=True, emit_split=True, keep_fp32=False, emit_stride=2)


def timeit(fn, name):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5):
        if i == 2:
            e0.record()
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"T={T} {name}: {e0.elapsed_time(e1) / 3:.3f} ms")
    return out


if which in ("fused", "both"):
    timeit(fused, "fused ingest (lsc MLP + conv_first*(lsc+1))")
    timeit(fused_lsc, "fused lsc MLP only")
if which in ("layers", "both"):
    timeit(layers, "layer by layer (4 + 1 launches)")
