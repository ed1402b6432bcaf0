"""Perf triage of the tcgen05 conv: split vs conv time, debug knobs (RCN_TC_DEBUG), stage count."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import _C, ops

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = 128
g = torch.Generator().manual_seed(0)
x = torch.randn(1, T, T, C, generator=g).to(dev)
w = (torch.randn(C, C, 3, 3, generator=g) / 34).to(dev)
b = torch.randn(C, generator=g).to(dev)
pc = ops.pack_weight(w, b)
y = torch.empty(1, T, T, C, device=dev)
hi = torch.empty(1, T, T, C, device=dev, dtype=torch.bfloat16)
lo = torch.empty_like(hi)
lib = _C.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ops._stream()

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def split(): _C.check(lib.rcn_split_bf16(P(x), C, T * T, C, C, 0, P(hi), P(lo), st))
d = _C.ConvDesc()
d.x, d.N, d.H, d.W, d.Cin, d.ldx = x.data_ptr(), 1, T, T, C, C
d.w, d.bias, d.k, d.stride, d.Cout = pc.w.data_ptr(), pc.bias.data_ptr(), 3, 1, C
d.y, d.ldy, d.store, d.act, d.slope, d.res_scale = y.data_ptr(), C, 0, 2, 0.01, 1.0
def conv(passes): _C.check(lib.rcn_conv2d_tc(ctypes.byref(d), P(hi), P(lo), P(pc.w_hi), P(pc.w_lo), C, passes, st))
flops = 2.0 * 9 * C * C * T * T
print(f"T={T} split {timeit(split):.3f} ms")
for dbg in (0, 1, 8, 2, 4, 6, 14):
    os.environ["RCN_TC_DEBUG"] = str(dbg)
    for passes in (3, 1):
        t = timeit(lambda: conv(passes))
        print(f"dbg={dbg:2d} passes={passes} conv {t:.3f} ms  {flops / t / 1e9:.1f} TF/s", flush=True)
os.environ["RCN_TC_DEBUG"] = "0"
for stages in (2, 3):
    os.environ["RCN_TC_STAGES"] = str(stages)
    print(f"stages={stages} passes=3 conv {timeit(lambda: conv(3)):.3f} ms")
print("tc_perf done")
