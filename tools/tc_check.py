"""tcgen05 conv engine vs the fp32 FFMA engine: numerics + timing on a list of layer shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
CASES = [  # N, H, W, Cin, Cout, k, extra
    (1, 8, 16, 64, 128, 1, {}),
    (1, 8, 16, 64, 64, 3, {}),
    (1, 16, 32, 128, 128, 3, {}),
    (2, 24, 40, 128, 128, 3, {"act": ops.ACT_LRELU, "slope": 0.01}),
    (1, 19, 27, 128, 320, 3, {}),
    (1, 16, 16, 320, 128, 1, {}),
    (1, 16, 16, 384, 224, 3, {"act": ops.ACT_GELU}),
    (1, 16, 16, 224, 128, 3, {}),
    (1, 32, 32, 48, 48, 3, {}),
    (1, 32, 32, 16, 32, 3, {}),
    (1, 32, 32, 128, 512, 3, {"store": ops.STORE_PS2}),
    (1, 32, 32, 128, 12, 3, {"store": ops.STORE_PS2_NCHW}),
    (1, 64, 64, 128, 128, 1, {"gdn": True}),
    (1, 256, 256, 128, 128, 3, {"res": True}),
    (1, 1024, 1024, 128, 128, 3, {}),
    (1, 12, 12, 320, 128, 3, {"stride": 2}),
    (1, 16, 16, 64, 64, 3, {"stride": 2}),
    (2, 64, 96, 128, 128, 3, {"stride": 2}),
    (1, 64, 64, 128, 128, 1, {"stride": 2}),
    (1, 1024, 1024, 128, 128, 1, {}),
    (1, 1024, 1024, 64, 64, 3, {}),
    (1, 1024, 1024, 64, 256, 1, {"act": ops.ACT_GELU}),
    (1, 1024, 1024, 256, 64, 1, {"res": True}),
]

def run(case, engine):
    N, H, W, Cin, Cout, k, ex = case
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    kw = {kk: v for kk, v in ex.items() if kk in ("act", "slope", "store", "stride")}
    if ex.get("gdn"):
        w = w.abs() * 0.1 + 0.1 * torch.eye(Cout, device=dev).reshape(Cout, Cin, 1, 1)
        b = b.abs() + 0.5
        kw.update(in_square=True, epi=ops.EPI_GDN, aux=x)
    if ex.get("res"):
        st = ex.get("stride", 1)
        kw["res"] = torch.randn(N, H // st, W // st, Cout, generator=g).to(dev)
    pc = ops.pack_weight(w, b)
    y = ops.conv2d(x, pc, engine=engine, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.conv2d(x, pc, engine=engine, out=y, **kw)
    e1.record()
    torch.cuda.synchronize()
    return y, e0.elapsed_time(e1) / 3

only = sys.argv[1:]
for i, case in enumerate(CASES):
    if only and str(i) not in only:
        continue
    ref, t0 = run(case, "fp32")
    line = f"case {i:2d} {case[:6]} fp32 {t0:8.3f} ms"
    for eng in ("bf16x3", "bf16"):
        try:
            y, t = run(case, eng)
            err = float((y - ref).abs().max() / ref.abs().max())
            flops = 2.0 * case[0] * case[1] * case[2] * case[3] * case[4] * case[5] ** 2 / case[6].get("stride", 1) ** 2
            line += f" | {eng}: rel {err:.2e} {t:8.3f} ms {flops / t / 1e9:8.1f} TF/s(incl split)"
        except Exception as e:
            line += f" | {eng}: ERROR {e}"
            break
    print(line, flush=True)
print("tc_check done")
