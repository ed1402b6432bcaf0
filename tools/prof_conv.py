"""Runs the dominant conv launch alone for `ncu --set full`: the full-resolution 128->128 3x3 convolution of the
g_s tail at T=2048 (the launch bench.py times for `roofline`).  usage: prof_conv.py [T] [fp32|bf16x3|bf16]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = (torch.randn(128, 128, 3, 3, generator=g) / 34.0).to(dev)
b = torch.randn(128, generator=g).to(dev)
pc = ops.pack_weight(w, b)
a = torch.randn(1, T, T, 128, device=dev)
o = torch.empty_like(a)
sp = ops.split_operand(a, 128, passes=3) if eng != "fp32" else None
for _ in range(4):
    ops.conv2d(a, pc, out=o, act=ops.ACT_LRELU, slope=0.01, engine=eng, presplit=sp)
torch.cuda.synchronize()
print("done")
