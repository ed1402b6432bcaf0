"""Per-call device time of EVERY C-ABI launch in one forward (CUDA events on the launching stream, no per-call sync).

  python tools/trace_step.py [T] [engine] [compress|forward]

Each rcn_* call is bracketed by two events; calls are grouped by (entry point, shape key) and printed with the
algorithmic bytes / FLOPs of the group so that the achieved GB/s and TFLOP/s of each layer shape can be read next to
its share of the step.  Perf triage only -- the numbers include launch gaps between the two events of one call only.
"""
import collections
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from realcamnet_b200 import synthetic as inputs, synthetic as weights
from realcamnet_b200 import _C, ops, raw2bit

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
mode = sys.argv[3] if len(sys.argv) > 3 else "forward"
ops.set_engine(eng)
dev = torch.device("cuda:0")
if mode == "gma":      # GMA_Block(dim 80) on a batch-32 T x T token map (BASELINE config 3)
    from realcamnet_b200 import groupmix
    m = groupmix.GMA_Block(80, 8)
    weights.fill_(m, seed=0)
    m = m.to(dev).eval()
    xg = torch.randn(32, T * T, 80, device=dev)
    run = lambda: m(xg, (T, T))
else:
    m = raw2bit.raw_compression_tcm_final()
    weights.fill_(m, seed=0)
    m = m.to(dev).eval()
    m.update()
    x = [t.to(dev) for t in inputs.make_inputs(T, seed=1234)]
    run = (lambda: m(x, emit_strings=True)) if mode == "forward" else (lambda: m.compress(x))
for _ in range(2):
    run()
torch.cuda.synchronize()

real = _C.lib()
records = []
SKIP = {"rcn_last_error", "rcn_launch_count", "rcn_version", "rcn_rans_encode", "rcn_rans_encode_packed", "rcn_pmf_to_quantized_cdf",
        "rcn_rans_decoder_create", "rcn_rans_decode", "rcn_rans_decoder_destroy", "rcn_groupmix_workspace_floats", "rcn_tc_prof"}


def conv_key(name, args):
    d = args[0]._obj
    passes = args[6] if name == "rcn_conv2d_tc" else 0
    Ho, Wo = d.H // d.stride, d.W // d.stride
    pix = d.N * Ho * Wo
    flop = 2.0 * d.k * d.k * d.Cin * d.Cout * pix
    if name == "rcn_conv2d_tc":
        cp = args[5]
        bin_ = d.N * d.H * d.W * cp * (4 if passes == 3 else 2)
    else:
        bin_ = d.N * d.H * d.W * d.Cin * 4
    bout = pix * d.Cout * ((4 if d.y else 0) + ((4 if d.y_lo else 2) if d.y_hi else 0))
    bout += pix * d.Cout * 4 * ((1 if d.res else 0) + (1 if d.epi else 0))
    key = (f"{name[4:]} {d.H}x{d.W} {d.Cin}->{d.Cout} k{d.k} s{d.stride} epi{d.epi} st{d.store} act{d.act}"
           f"{' res' if d.res else ''}{' cs' if d.cscale else ''}{' y' if d.y else ''}{' planes' if d.y_hi else ''}"
           f"{' xsplit' if not d.x else ''}")
    return key, bin_ + bout, flop


def generic_key(name, args):
    ints = [a for a in args if isinstance(a, int) and not isinstance(a, bool)]
    return f"{name[4:]} {ints[:6]}", 0, 0


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if name in SKIP:
            return fn

        def wrapped(*args):
            try:
                key, nbytes, flop = conv_key(name, args) if name in ("rcn_conv2d", "rcn_conv2d_tc") else generic_key(name, args)
            except Exception:
                key, nbytes, flop = name, 0, 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            records.append((key, nbytes, flop, e0, e1))
            return rc

        return wrapped


_C._lib = Proxy()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
with torch.no_grad():
    run()
s1.record()
torch.cuda.synchronize()
_C._lib = real
total = s0.elapsed_time(s1)
agg = collections.OrderedDict()
for key, nb, fl, e0, e1 in records:
    a = agg.setdefault(key, [0, 0.0, 0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += nb
    a[3] += fl
ksum = sum(a[1] for a in agg.values())
print(f"T={T} engine={eng} mode={mode}: step {total:.2f} ms on the stream; {len(records)} calls, {ksum:.2f} ms inside calls")
print(f"{'ms':>8s} {'share':>6s} {'n':>4s} {'ms/call':>8s} {'GB/s':>7s} {'TF/s':>7s}  call")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("RCN_TRACE_ROWS", "70"))]:
    gbs = a[2] / a[1] / 1e6 if a[2] else 0.0
    tfs = a[3] / a[1] / 1e9 if a[3] else 0.0
    print(f"{a[1]:8.3f} {100 * a[1] / total:5.1f}% {a[0]:4d} {a[1] / a[0]:8.4f} {gbs:7.0f} {tfs:7.1f}  {key}")
by_fn = collections.Counter()
for key, a in agg.items():
    by_fn[key.split()[0]] += a[1]
print("by entry point:", ", ".join(f"{k} {v:.2f}" for k, v in by_fn.most_common()))
