"""Perf triage of the tcgen05 conv on arbitrary layer shapes: python tools/tc_perf2.py H Cin Cout k [passes]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import _C, ops

dev = torch.device("cuda:0")
lib = _C.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(T, Cin, Cout, k, res=False, dbgs=(0, 1, 8, 2, 4)):
    g = torch.Generator().manual_seed(0)
    Cp = (Cin + 63) // 64 * 64
    x = torch.randn(1, T, T, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    pc = ops.pack_weight(w, b)
    y = torch.empty(1, T, T, Cout, device=dev)
    r = torch.randn(1, T, T, Cout, device=dev) if res else None
    hi = torch.empty(1, T, T, Cp, device=dev, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    st = ops._stream()
    _C.check(lib.rcn_split_bf16(P(x), Cin, T * T, Cin, Cp, 0, P(hi), P(lo), st))
    d = _C.ConvDesc()
    d.x, d.N, d.H, d.W, d.Cin, d.ldx = x.data_ptr(), 1, T, T, Cin, Cin
    d.w, d.bias, d.k, d.stride, d.Cout = pc.w.data_ptr(), pc.bias.data_ptr(), k, 1, Cout
    d.y, d.ldy, d.store, d.act, d.slope, d.res_scale = y.data_ptr(), Cout, 0, 2, 0.01, 1.0
    if res:
        d.res, d.ldres = r.data_ptr(), Cout
    flops = 2.0 * k * k * Cin * Cout * T * T
    obytes = T * T * Cout * 4 * (2 if res else 1) + T * T * Cp * 4
    for passes in (3, 1):
        for dbg in dbgs:
            os.environ["RCN_TC_DEBUG"] = str(dbg)
            t = timeit(lambda: _C.check(lib.rcn_conv2d_tc(ctypes.byref(d), P(hi), P(lo), P(pc.w_hi), P(pc.w_lo), Cp, passes, st)))
            print(f"T={T} {Cin}->{Cout} k{k} res={int(res)} passes={passes} dbg={dbg:2d}: {t:.3f} ms  {flops * passes / t / 1e9:7.1f} TF/s issued  "
                  f"{obytes / t / 1e6:7.1f} GB/s", flush=True)
    os.environ["RCN_TC_DEBUG"] = "0"


if __name__ == "__main__":
    if len(sys.argv) > 4:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        run(1024, 64, 256, 1)
        run(2048, 128, 128, 1)
        run(1024, 128, 128, 1, res=True)
        run(2048, 128, 128, 3)
        run(2048, 128, 128, 3, res=True)
        run(1024, 64, 64, 3)
        run(128, 128, 128, 3, dbgs=(0,))
        run(128, 64, 64, 3, dbgs=(0,))
