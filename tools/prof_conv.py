"""Runs one conv launch shape alone for `ncu --set full`.  Default: the full-resolution 128->128 3x3 convolution of the
g_s tail at T=2048 (the launch bench.py times for `roofline`).
usage: prof_conv.py [T] [fp32|bf16x3|bf16|fp16] [Cin] [Cout] [k] [act] [emit_planes 0|1] [res 0|1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
Cin = int(sys.argv[3]) if len(sys.argv) > 3 else 128
Cout = int(sys.argv[4]) if len(sys.argv) > 4 else 128
k = int(sys.argv[5]) if len(sys.argv) > 5 else 3
act = int(sys.argv[6]) if len(sys.argv) > 6 else ops.ACT_LRELU
emit = bool(int(sys.argv[7])) if len(sys.argv) > 7 else False
res = bool(int(sys.argv[8])) if len(sys.argv) > 8 else False
ops.set_engine(eng)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
b = torch.randn(Cout, generator=g).to(dev)
pc = ops.pack_weight(w, b)
a = torch.randn(1, T, T, Cin, device=dev)
r = torch.randn(1, T, T, Cout, device=dev) if res else None
sp = ops.split_operand(a, pc.cp) if eng != "fp32" else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    if i == 2:
        e0.record()
    ops.conv2d(a, pc, act=act, slope=0.01, engine=eng, presplit=sp, emit_split=emit, keep_fp32=not emit, res=r)
e1.record()
torch.cuda.synchronize()
print(f"T={T} {eng} {Cin}->{Cout} k{k} act{act} emit={emit} res={res}: {e0.elapsed_time(e1) / 3:.3f} ms/launch")
