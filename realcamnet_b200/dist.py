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


def my_tiles(num_tiles: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: tile t -> rank t mod world."""
    return list(range(rank, num_tiles, world))


def scatter_tiles(tiles: torch.Tensor | None, num_tiles: int, tile_shape: Sequence[int], src: int = 0, device=None,
                  dtype=torch.float32) -> torch.Tensor:
    """Rank `src` holds tiles (T,C,h,w); every rank receives the (len(my_tiles),C,h,w) stack it owns."""
    rank, world = dist.get_rank(), dist.get_world_size()
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
    rank, world = dist.get_rank(), dist.get_world_size()
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
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = scatter_tiles(tiles_on_src, num_tiles, tile_shape, device=device)
    streams = [codec(mine[i:i + 1], t) for i, t in enumerate(my_tiles(num_tiles, rank, world))]
    return gather_bitstreams(streams, num_tiles, device=device)
