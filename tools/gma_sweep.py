"""BASELINE config 3: the GroupMix block (GMA_Block) isolated, batch 32, token-map sweep, achieved HBM GB/s.

  python tools/gma_sweep.py [out.json]

Algorithmic bytes per token: x is read twice (the softmax over all N tokens / k^T v needs a reduction pass before
the apply pass) and the result written once = 3*C values; the B200 path keeps activations in fp32, so 12*C bytes per token
(960 B at C=80, 2400 B at C=200).  Time: CUDA events around one GMA_Block forward (3 warm-ups, median of 5), inputs resident.
The 1024x1024 map runs at batch 8 (batch 32 would need > 130 GB of fp32 intermediates), the smaller maps at batch 32.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from realcamnet_b200 import synthetic as weights
from realcamnet_b200 import groupmix, ops

dev = torch.device("cuda:0")
peak = 6540.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for dim in (80, 200):
    m = groupmix.GMA_Block(dim, 8)
    weights.fill_(m, seed=0)
    m = m.to(dev).eval()
    for side, B in ((128, 32), (256, 32), (512, 32), (1024, 8)):
        need = B * side * side * dim * 4 * 14          # x + ~13x of intermediates
        free = torch.cuda.mem_get_info()[0]
        if need > 0.5 * free:
            rows.append({"dim": dim, "map": side, "batch": B, "skipped": f"needs ~{need / 2**30:.0f} GiB"})
            continue
        x = torch.randn(B, side * side, dim, device=dev)
        ts = []
        with torch.no_grad():
            for i in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = m(x, (side, side))
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
                del y
        ms = sorted(ts)[len(ts) // 2]
        tok = B * side * side
        gbs = tok * 12 * dim / (ms * 1e-3) / 1e9
        rows.append({"dim": dim, "map": side, "batch": B, "tokens": tok, "ms": ms, "tokens_per_s": tok / (ms * 1e-3),
                     "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak})
        print(f"GMA_Block dim={dim:3d} map={side:4d}^2 batch={B:2d}: {ms:8.3f} ms  {tok / ms / 1e3:8.1f} Mtok/s  "
              f"{gbs:7.1f} GB/s algorithmic ({100 * gbs / peak:4.1f}% of {peak:.0f})", flush=True)
        del x
        torch.cuda.empty_cache()
out = {"config": "BASELINE configs[2]: GMA_Block isolated, fp32 activations, engine " + ops.get_engine(), "hbm_peak_GBps": peak, "rows": rows}
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
