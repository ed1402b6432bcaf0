"""Run-to-run determinism of the tcgen05 engine and forward-vs-decompress consistency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import synthetic as inputs, synthetic as weights
from realcamnet_b200 import ops, raw2bit

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (H, W, Cin, Cout, k) in [(64, 64, 128, 128, 3), (16, 16, 384, 224, 3), (4, 4, 192, 512, 3), (128, 128, 128, 12, 3), (32, 32, 128, 128, 1)]:
    x = torch.randn(1, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    pc = ops.pack_weight(w, None)
    outs = [ops.conv2d(x, pc, engine="bf16x3").clone() for _ in range(4)]
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(f"conv {H}x{W} {Cin}->{Cout} k{k}: 4 runs bitwise equal: {same}", flush=True)

for eng in ("fp32", "bf16x3"):
    ops.set_engine(eng)
    m = raw2bit.raw_compression_tcm_final()
    weights.fill_(m, seed=0)
    m = m.to(dev).eval(); m.update()
    x = [t.to(dev) for t in inputs.make_inputs(256, seed=1234)]
    o1 = m(x, emit_strings=True); o2 = m(x, emit_strings=True)
    print(eng, "forward twice: x_hat equal", torch.equal(o1["x_hat"], o2["x_hat"]), "y equal", torch.equal(o1["y"], o2["y"]),
          "bytes equal", o1["strings"] == o2["strings"], flush=True)
    d = m.decompress(o1["strings"], o1["shape"])
    a, b = d["x_hat"], o1["x_hat"].clamp(0, 1)
    diff = (a - b).abs()
    print(eng, "decompress vs forward: equal", torch.equal(a, b), "max abs diff", float(diff.max()), "n differing", int((diff > 0).sum()),
          "of", a.numel(), flush=True)
    d2 = m.decompress(o1["strings"], o1["shape"])
    print(eng, "decompress twice equal", torch.equal(d["x_hat"], d2["x_hat"]), flush=True)
print("done")
