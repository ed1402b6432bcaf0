"""Times the Swin MLP of one Block (C = 64, hidden = 256) on a T x T token map: the fused kernel (rcn_mlp_fused) against the two
conv launches.  usage: prof_mlp.py [T] [fused|layers|both]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from realcamnet_b200 import ops, synthetic
from realcamnet_b200.layers import Linear

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "both"
dev = torch.device("cuda:0")
ops.set_engine("bf16x3")
fc1, fc2 = Linear(64, 256), Linear(256, 64)
synthetic.fill_(fc1, seed=11)
synthetic.fill_(fc2, seed=12)
fc1, fc2 = fc1.to(dev), fc2.to(dev)
t = torch.randn(1, T, T, 64, device=dev)
both = torch.randn(1, T, T, 128, device=dev)
res = both[..., 64:]
tsp = ops.split_operand(t, 64)
csp = ops.alloc_planes(1, T, T, 128, dev)


def fused():
    return ops.mlp_fused(tsp, fc1, fc2, res=res, split_out=csp.channels(64, 128), keep_fp32=False)


def layers():
    h, hsp = fc1._f(t, act=ops.ACT_GELU, emit_split=True, keep_fp32=False, presplit=tsp)
    return fc2._f(h, res=res, presplit=hsp, split_out=csp.channels(64, 128), keep_fp32=False)


def timeit(fn, name):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5):
        if i == 2:
            e0.record()
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    gb = T * T * (64 * 4 + 64 * 4 + 64 * 4) / 1e9      # LN planes in (hi + lo), residual in, operand planes out (hi + lo)
    print(f"T={T} {name}: {ms:.3f} ms  ({gb / ms * 1e3:.0f} GB/s on the {gb:.2f} GB of compulsory traffic)")


if which in ("fused", "both"):
    timeit(fused, "fused Swin MLP (fc1 + GELU + fc2 + residual, planes out)")
if which in ("layers", "both"):
    timeit(layers, "two launches (fc1 + GELU -> hidden planes -> fc2 + residual)")
