"""Section timing of raw_compression_tcm_final.forward at tile T: GPU time (sync'd) and CPU launch time per section."""
import os, sys, time
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
for _ in range(2): m(x, emit_strings=True)
torch.cuda.synchronize()

class Sec:
    def __init__(self): self.rows = []
    def run(self, name, fn):
        torch.cuda.synchronize(); n0 = ops.launch_count(); t0 = time.perf_counter()
        r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        self.rows.append((name, (t1 - t0) * 1e3, (t2 - t0) * 1e3, ops.launch_count() - n0)); return r
S = Sec()
with torch.no_grad():
    y, lsc, local = S.run("analysis (g_a + conditioning)", lambda: m._analysis(x))
    z = S.run("h_a", lambda: m._h_a(y))
    z_hat, z_lik, z_sym = S.run("entropy bottleneck", lambda: m.entropy_bottleneck._f(z, want_symbols=True))
    ms, ss, h, w = S.run("h_mean_s + h_scale_s", lambda: m._alloc_supports(z_hat))
    N, sl = 1, 64
    table = m._scale_table_dev(); gc = m.gaussian_conditional
    sym = torch.empty((5, N * sl * h * w), device=dev, dtype=torch.int32); idx = torch.empty_like(sym)
    y_lik = ops.empty(N, h, w, 320, like=y)
    def slices():
        for i in range(5):
            lrp_sup, cin, mu, scale = m._slice_params(i, ms, ss)
            ops.gaussian_conditional(y[..., sl * i: sl * (i + 1)], mu, scale, table, y_hat=lrp_sup[..., cin:], lik=y_lik[..., sl*i:sl*(i+1)],
                                     symbols=sym[i], indexes=idx[i], scale_bound=gc._scale_bound, lik_bound=gc.likelihood_bound)
            m._finish_slice(i, lrp_sup, cin, ms, ss)
    S.run("5-slice entropy-parameter loop", slices)
    S.run("g_s (synthesis)", lambda: m._g_s(ms[..., 320:]))
    S.run("NCHW conversions of outputs", lambda: [ops.to_nchw(t) for t in (y, local[2], lsc, y_lik, z_lik)])
    S.run("D2H symbols + host rANS", lambda: (m._encode_y(sym, idx), m.entropy_bottleneck.compress_symbols(z_sym)))
print(f"T={T} engine={eng}")
print(f"{'section':38s} {'cpu-launch ms':>14s} {'gpu-done ms':>12s} {'launches':>9s}")
for r in S.rows: print(f"{r[0]:38s} {r[1]:14.2f} {r[2]:12.2f} {r[3]:9d}")
print(f"{'SUM':38s} {sum(r[1] for r in S.rows):14.2f} {sum(r[2] for r in S.rows):12.2f} {sum(r[3] for r in S.rows):9d}")
t0 = time.perf_counter(); o = m(x, emit_strings=True); torch.cuda.synchronize(); print("full forward(emit_strings) wall ms", (time.perf_counter() - t0) * 1e3)
