"""GPU throughput of the ISP variants (BASELINE config 1: one 4x256x256 packed-Bayer tile through LiteISPNet_GFM_LSC).

  python tools/liteisp_bench.py [T=256] [T2=1024]   -> one JSON line per (model, tile): sensor MP/s, ms per tile

Inputs resident on the device, eager launches and CUDA-graph replay (enable_cuda_graphs), CUDA events, median of 7 after 3 warm-ups.  1 MP = 1e6 sensor photosites (4*T*T per tile).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from realcamnet_b200 import LiteISP, ops, synthetic

dev = torch.device("cuda:0")
tiles = [int(a) for a in sys.argv[1:]] or [256, 1024]
for name in ("LiteISPNet_GFM_LSC", "LiteISPNet", "ISPUNet_GFM_LSC", "ResUNet", "MWISP"):
    m = getattr(LiteISP, name)()
    synthetic.fill_(m, seed=0)
    m = m.to(dev).eval()
    for T, graphs in [(T, g) for T in tiles for g in (False, True)]:
        m.enable_cuda_graphs(graphs)
        x = [t.to(dev) for t in synthetic.make_inputs(T, seed=1235)]
        with torch.no_grad():
            for _ in range(3):
                m(x)
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = ops.launch_count()
                e0.record()
                m(x)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                launches = ops.launch_count() - n0
        ms = sorted(ts)[len(ts) // 2]
        print(json.dumps({"model": name, "tile": T, "ms_per_tile": ms, "sensor_mp_per_s": 4.0 * T * T / 1e6 / (ms / 1e3),
                          "launches": launches, "engine": ops.get_engine(), "cuda_graph": graphs}), flush=True)
