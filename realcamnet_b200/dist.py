"""Multi-GPU tile sharding for the RAW->bitstream path (one process per GPU, torch.distributed).

The path shards by independent units (tiles): tile t goes to rank t mod G, weights are replicated, no data-path
collective exists inside a tile.  The only exchanges are the frame scatter in and the variable-length bitstream
gather out; both are plain torch.distributed collectives (NCCL over NVLink on the GPU box,
gloo in the CPU tests) -- payloads are a few MB per frame, far from any link bound, so no fused kernel is warranted.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def _rank_world():
    """(rank, world) of the default process group; (0, 1) when torch.distributed is not initialised (single GPU)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_tiles(num_tiles: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: tile t -> rank t mod world."""
    return list(range(rank, num_tiles, world))


def scatter_tiles(tiles: torch.Tensor | None, num_tiles: int, tile_shape: Sequence[int], src: int = 0, device=None,
                  dtype=torch.float32) -> torch.Tensor:
    """Rank `src` holds tiles (T,C,h,w); every rank receives the (len(my_tiles),C,h,w) stack it owns."""
    rank, world = _rank_world()
    mine = my_tiles(num_tiles, rank, world)
    out = torch.empty((len(mine), *tile_shape), device=device, dtype=dtype)
    if world == 1:
        out.copy_(tiles[mine])
        return out
    # point-to-point sends keep the traffic at exactly one copy of each tile (a broadcast would move the whole frame G times)
    ops = []
    if rank == src:
        for r in range(world):
            idx = my_tiles(num_tiles, r, world)
            if r == src:
                out.copy_(tiles[idx].to(out.device))
            elif idx:
                ops.append(dist.P2POp(dist.isend, tiles[idx].contiguous().to(out.device), r))
    elif mine:
        ops.append(dist.P2POp(dist.irecv, out, src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def gather_bitstreams(local: List[bytes], num_tiles: int, dst: int = 0, device=None) -> List[bytes] | None:
    """Variable-length gather: each rank contributes the byte strings of its tiles (in my_tiles order);
    rank `dst` returns the list for all tiles in tile order, the others return None."""
    rank, world = _rank_world()
    if world == 1:
        return list(local)
    device = device or torch.device("cpu")
    per_rank = (num_tiles + world - 1) // world
    lens = torch.zeros((per_rank,), dtype=torch.int64, device=device)
    for i, b in enumerate(local):
        lens[i] = len(b)
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    cap = int(max(int(l.sum()) for l in all_lens))
    buf = torch.zeros((max(cap, 1),), dtype=torch.uint8, device=device)
    blob = b"".join(local)
    if blob:
        buf[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst)
    if rank != dst:
        return None
    out: List[bytes] = [b""] * num_tiles
    for r in range(world):
        data = gathered[r].cpu().numpy().tobytes()
        off = 0
        for i, t in enumerate(my_tiles(num_tiles, r, world)):
            n = int(all_lens[r][i])
            out[t] = data[off:off + n]
            off += n
    return out


def compress_frame_sharded(tiles_on_src: torch.Tensor | None, num_tiles: int, tile_shape: Sequence[int],
                           codec: Callable[[torch.Tensor, int], bytes], device=None) -> List[bytes] | None:
    """scatter -> per-rank codec(tile, tile_index) -> gather; `codec` is the per-tile RAW->bitstream function."""
    rank, world = _rank_world()
    mine = scatter_tiles(tiles_on_src, num_tiles, tile_shape, device=device)
    streams = [codec(mine[i:i + 1], t) for i, t in enumerate(my_tiles(num_tiles, rank, world))]
    return gather_bitstreams(streams, num_tiles, device=device)


def compress_frame_distributed(model, frame: torch.Tensor | None, height: int, width: int, tile: int, device, model_id: int = 0,
                               max_batch: int = 0) -> bytes | None:
    """BASELINE config 4: one packed-Bayer frame (4,height,width), held by rank 0, -> RCNB container bytes on rank 0.

    Rank 0 uploads the frame, tiles it on the device and scatters tile t to rank t mod G (point-to-point); the frame-level colour
    condition is broadcast once; every rank pushes ITS tiles through frame.compress_tiles (equal batches, host coder overlapped
    with the next batch); the variable-length streams are gathered on rank 0 and packed.  The container does not depend on G."""
    from . import container, tiler
    from . import frame as rframe

    import os
    import time

    timing = bool(os.environ.get("RCN_FRAME_TIMING"))
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    rank, world = _rank_world()
    ny, nx = tiler.tile_grid(height, width, tile)
    meta, ntiles = (height, width, ny, nx), ny * nx
    mark("start")
    if rank == 0:
        fdev = frame.to(device, non_blocking=True)
        tiles = tiler.split_frame(fdev, tile)[0]
        cond = rframe.frame_condition(fdev)
    else:
        tiles, cond = None, torch.empty((1, 4, 256, 256), device=device)
    mark("upload+tile+cond")
    if world > 1:
        dist.broadcast(cond, 0)
    mine = my_tiles(ntiles, rank, world)
    local = scatter_tiles(tiles, ntiles, (4, tile, tile), device=device)
    mark("scatter")
    coords = _coords_cached(meta, tile, tuple(mine), device)
    mark("coords")
    out = rframe.compress_tiles(model, local, cond, coords, max_batch=max_batch)
    mark("compress_tiles")
    all_y = gather_bitstreams([o[0] for o in out], ntiles, device=device)
    all_z = gather_bitstreams([o[1] for o in out], ntiles, device=device)
    mark("gather")
    if rank != 0:
        return None
    shape = out[0][2]
    recs = [container.TileStreams(t, shape, all_y[t], all_z[t]) for t in range(ntiles)]
    blob = container.pack(container.FrameHeader(model_id, height, width, tile, ny, nx, ntiles), recs)
    mark("pack")
    if timing:
        import sys
        print("frame timing (ms): " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f}" for a, b in zip(marks, marks[1:])), file=sys.stderr)
    return blob


_coords_cache: dict = {}


def _coords_cached(meta, tile, mine, device):
    """Per-tile coordinate maps depend on the frame geometry only: built once per (geometry, tile share, device)."""
    from . import tiler

    key = (meta, tile, mine, str(device))
    c = _coords_cache.get(key)
    if c is None:
        if len(_coords_cache) > 8:
            _coords_cache.clear()
        c = _coords_cache[key] = tiler.tiles_coords(meta, tile, mine, device=device)
    return c
