"""Per-layer-shape time of every conv2d call in one forward (sync'd CUDA events around split+conv)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realcamnet_b200 import synthetic as inputs, synthetic as weights
from realcamnet_b200 import ops, raw2bit

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
ops.set_engine(eng)
dev = torch.device("cuda:0")
m = raw2bit.raw_compression_tcm_final(); weights.fill_(m, seed=0); m = m.to(dev).eval(); m.update()
x = [t.to(dev) for t in inputs.make_inputs(T, seed=1234)]
for _ in range(2): m(x)   # (superseded by tools/trace_step.py, which also covers the non-conv kernels)
torch.cuda.synchronize()
stats = collections.OrderedDict()
orig = ops.conv2d
def timed(x_, pc, stride=1, **kw):
    N, H, W, C = x_.shape if x_ is not None else kw['presplit'].key[1:5]   # planes-only input
    key = (H, W, C, pc.cout, pc.k, stride, kw.get("epi", 0), kw.get("store", 0), "res" if kw.get("res") is not None else "", "pre" if kw.get("presplit") is not None else "")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(x_, pc, stride=stride, **kw); e1.record(); e1.synchronize()
    s = stats.setdefault(key, [0, 0.0]); s[0] += 1; s[1] += e0.elapsed_time(e1)
    return r
ops.conv2d = timed
import realcamnet_b200.layers as L
with torch.no_grad(): m(x)
ops.conv2d = orig
tot = sum(v[1] for v in stats.values())
print(f"T={T} engine={eng}: {sum(v[0] for v in stats.values())} conv calls, {tot:.1f} ms total (incl. split, sync'd per call)")
print(f"{'H':>5s} {'W':>5s} {'Cin':>4s} {'Cout':>4s} k s epi st res pre {'n':>4s} {'ms':>8s} {'ms/call':>8s} {'TF/s':>7s}")
for k, v in sorted(stats.items(), key=lambda kv: -kv[1][1])[:45]:
    H, W, C, Co, kk, s, epi, st, res, pre = k
    fl = 2.0 * (H // s) * (W // s) * C * Co * kk * kk * v[0]
    print(f"{H:5d} {W:5d} {C:4d} {Co:4d} {kk} {s} {epi:3d} {st:2d} {res:3s} {pre:3s} {v[0]:4d} {v[1]:8.2f} {v[1]/v[0]:8.3f} {fl/v[1]/1e9:7.1f}")
