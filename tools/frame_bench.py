"""BASELINE config 4: a 3840x2160 RAW frame (packed 4 x 1080 x 1920... see below) tiled 512x512, tile-sharded across the ranks.

  python tools/frame_bench.py [--tile 512] [--reps 3] [--out file.json]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/frame_bench.py   (N GPUs, NCCL)

Frame and tiles live in the same (packed) space, SURVEY.md 8d "secondary interpretation": a packed 4 x 2160 x 3840 frame is padded to
2560 x 4096 = 5 x 8 = 40 tiles of 4 x 512 x 512, tile t -> rank t mod G; rank 0 scatters the tiles (point-to-point over NCCL), every
rank runs raw_compression_tcm_final.compress on its tiles, the variable-length bitstreams are gathered on rank 0 and packed into one
RCNB container.  Timed per frame with CUDA events + barriers (max over ranks); sensor megapixels = 4 * 2160 * 3840 / 1e6.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from realcamnet_b200 import synthetic as weights
from realcamnet_b200 import container, frame, raw2bit, tiler
from realcamnet_b200 import dist as rdist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    m = raw2bit.raw_compression_tcm_final()
    weights.fill_(m, seed=0)
    m = m.to(dev).eval()
    m.update()
    m.enable_cuda_graphs(True)
    T = args.tile
    g = torch.Generator().manual_seed(99)
    fr = torch.rand(4, args.height, args.width, generator=g).pin_memory() if rank == 0 else None
    meta = (args.height, args.width) + tiler.tile_grid(args.height, args.width, T)
    H, W, ny, nx = meta
    ntiles = ny * nx
    cond = frame.frame_condition(fr).to(dev) if rank == 0 else torch.empty(1, 4, 256, 256, device=dev)
    if world > 1:
        dist.broadcast(cond, 0)
    mine = rdist.my_tiles(ntiles, rank, world)

    def one_frame():
        # the maintained path (what bench.py's `frame4k` times): H2D of the frame, tiling on the device, NCCL scatter, batched graph
        # replays with the host range coder overlapped, bitstream gather, RCNB container on rank 0
        return rdist.compress_frame_distributed(m, fr, args.height, args.width, T, dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blob = one_frame()           # warm-up: weight packing, graph capture
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        blob2 = one_frame()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.reps
    wall = (time.perf_counter() - t0) / args.reps * 1e3
    tm = torch.tensor([ms, wall], device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert blob2 == blob, "frame bitstream is not deterministic"
        hdr, recs = container.unpack(blob)
        assert len(recs) == ntiles
        mp = 4.0 * args.height * args.width / 1e6
        line = {"config": f"BASELINE configs[3]: packed 4x{args.height}x{args.width} RAW frame, {ntiles} tiles of 4x{T}x{T}, tile t -> rank t mod {world}, "
                          "compress() per tile + scatter/gather + RCNB container", "n_gpus": world, "tiles": ntiles,
                "ms_per_frame_device": float(tm[0]), "ms_per_frame_wall": float(tm[1]), "MP_per_s": mp / (float(tm[1]) * 1e-3),
                "frames_per_s": 1e3 / float(tm[1]), "container_bytes": len(blob), "bits_per_sensor_pixel": 8.0 * len(blob) / (mp * 1e6)}
        print(json.dumps(line))
        if args.out:
            json.dump(line, open(args.out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
