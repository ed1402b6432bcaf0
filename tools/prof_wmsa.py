"""One window-attention launch for ncu: python tools/prof_wmsa.py [T] [C] [head_dim] [ws] [shifted]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hd = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ws = int(sys.argv[4]) if len(sys.argv) > 4 else 8
sh = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
dev = torch.device("cuda:0")
qkv = torch.randn(1, T, T, 3 * C, device=dev)
rel = torch.randn(C // hd, 2 * ws - 1, 2 * ws - 1, device=dev) * 0.02
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 2:
        e0.record()
    o, sp = ops.wmsa(qkv, rel, hd, ws, sh, emit_split=True)
e1.record()
torch.cuda.synchronize()
print(f"wmsa T={T} C={C} hd={hd} ws={ws} shifted={sh}: {e0.elapsed_time(e1) / 4:.3f} ms/launch")
