"""Precision policy table: for each engine assignment of the synthesis transform, the error of x_hat against the
CPU oracle on the same input, the decode == forward check, and the step time at the bench tile.

  python tools/precision_table.py [T_err=512] [T_time=2048]      -> markdown table on stdout

Rows: the global engine is bf16x3 everywhere a quantisation decision depends on the result; only `g_s[tail_start:]` changes.
The oracle is used as the checker (this tool is test infrastructure, like tests/).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import refpath
from realcamnet_b200 import ops, raw2bit, synthetic

T_err = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T_time = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
ops.set_engine("bf16x3")
m = raw2bit.raw_compression_tcm_final()
synthetic.fill_(m, seed=0)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.to(dev).eval()
m.update()
x = synthetic.make_inputs(T_err, seed=1234)
xd = [t.to(dev) for t in x]
xt = [t.to(dev) for t in synthetic.make_inputs(T_time, seed=1234)]
torch.set_num_threads(min(32, os.cpu_count() or 1))
ref = refpath.final_forward(sd, x)
ref_sym = torch.round(ref["para"]["y"] - ref["para"]["means"])

rows = [("bf16x3 everywhere (policy off)", "bf16x3", None), ("fp16 x1: tail g_s[9:] (DEFAULT)", "fp16", None), ("fp16 x1: g_s[6:]", "fp16", 6),
        ("fp16 x1: g_s[3:]", "fp16", 3), ("fp16 x1: all of g_s", "fp16", 0), ("bf16 x1: tail g_s[9:]", "bf16", None), ("bf16 x1: all of g_s", "bf16", 0)]
print(f"| synthesis engine assignment | x_hat max err / max vs bf16x3 synthesis (T={T_err}) | PSNR vs bf16x3 synthesis | x_hat max err / max vs oracle "
      f"| PSNR vs oracle | symbols differing from the oracle's | decode == forward | ms / step (T={T_time}) |")
print("|---|---:|---:|---:|---:|---:|---|---:|")
base = None
for name, te, ts in rows:
    m.tail_engine, m.tail_start = te, ts
    m.enable_cuda_graphs(False)
    out = m(xd, emit_strings=True)
    xh = out["x_hat"].cpu()
    err = float((xh - ref["x_hat"]).abs().max() / ref["x_hat"].abs().max())
    psnr = refpath.psnr(xh, ref["x_hat"])
    if base is None:
        base = xh      # same y_hat in every row (the analysis / entropy stages never change engine): isolates the synthesis engine
    berr = float((xh - base).abs().max() / base.abs().max())
    bpsnr = refpath.psnr(xh, base) if berr > 0 else float("inf")
    sym = torch.round(out["para"]["y"] - out["para"]["means"]).cpu()
    nmis = int((sym != ref_sym).sum())
    d = m.decompress(out["strings"], out["shape"])
    same = bool(torch.equal(d["x_hat"], out["x_hat"].clamp(0, 1)))
    del out, d
    m.enable_cuda_graphs(True)
    for _ in range(3):
        m(xt, emit_strings=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(xt, emit_strings=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    m.enable_cuda_graphs(False)      # drops the captured graphs (and their memory pool) before the next row
    torch.cuda.empty_cache()
    print(f"| {name} | {berr:.2e} | {bpsnr:.1f} dB | {err:.2e} | {psnr:.1f} dB | {nmis} / {sym.numel()} | {same} | {ms:.2f} |", flush=True)
