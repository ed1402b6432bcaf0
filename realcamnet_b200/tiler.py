"""Frame tiler for the RAW codec: packed-Bayer frame -> independent zero-padded tiles (+ per-tile coordinate maps)
and back.  Tiles are independent images for every layer of the path (all global reductions -- channel attention
pooling, InstanceNorm, k-softmax -- are per image), so no halo is exchanged; padding is bottom/right with zeros,
the policy of the reference's ``pad_to_multiple_of_16`` (models/LiteISP.py:84-105).  Host-side index arithmetic only.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


def tile_grid(H: int, W: int, tile: int) -> Tuple[int, int]:
    return math.ceil(H / tile), math.ceil(W / tile)


def split_frame(frame: torch.Tensor, tile: int) -> Tuple[torch.Tensor, Tuple[int, int, int, int]]:
    """frame (C,H,W) or (1,C,H,W) -> tiles (T,C,tile,tile), row-major over the tile grid; returns (tiles, (H,W,ny,nx))."""
    if frame.dim() == 4:
        frame = frame[0]
    C, H, W = frame.shape
    ny, nx = tile_grid(H, W, tile)
    padded = frame.new_zeros((C, ny * tile, nx * tile))
    padded[:, :H, :W] = frame
    tiles = padded.reshape(C, ny, tile, nx, tile).permute(1, 3, 0, 2, 4).reshape(ny * nx, C, tile, tile)
    return tiles.contiguous(), (H, W, ny, nx)


def tile_coords(meta: Tuple[int, int, int, int], tile: int, index: int, device=None) -> torch.Tensor:
    """(1,2,tile,tile) normalised pixel-centre coordinates of tile `index` inside its frame:
    channel 0 = x in [-1,1] along W, channel 1 = y in [-1,1] along H, normalised by the REAL
    (unpadded) frame size -- like the reference's inputs, which span [-1,1] over the actual image -- so the optical centre stays
    at (0,0) whatever the tile size; zero-padded pixels simply fall outside [-1,1].  (RCNB container version 2 convention.)"""
    H, W, ny, nx = meta
    ty, tx = divmod(index, nx)
    ys = (torch.arange(tile, device=device, dtype=torch.float32) + ty * tile) / max(H - 1, 1) * 2 - 1
    xs = (torch.arange(tile, device=device, dtype=torch.float32) + tx * tile) / max(W - 1, 1) * 2 - 1
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([xx, yy])[None].contiguous()


def tiles_coords(meta: Tuple[int, int, int, int], tile: int, indices, device=None) -> torch.Tensor:
    """(len(indices),2,tile,tile): tile_coords of several tiles in one tensor (same values, bit for bit)."""
    H, W, ny, nx = meta
    idx = torch.as_tensor(list(indices), dtype=torch.int64, device=device)
    ty, tx = idx // nx, idx % nx
    ar = torch.arange(tile, device=device, dtype=torch.float32)
    ys = (ar[None, :] + (ty * tile).to(torch.float32)[:, None]) / max(H - 1, 1) * 2 - 1        # (n, tile)
    xs = (ar[None, :] + (tx * tile).to(torch.float32)[:, None]) / max(W - 1, 1) * 2 - 1
    n = idx.numel()
    return torch.stack([xs[:, None, :].expand(n, tile, tile), ys[:, :, None].expand(n, tile, tile)], dim=1).contiguous()


def stitch(tiles: List[torch.Tensor], meta: Tuple[int, int, int, int], tile: int, scale: int = 2) -> torch.Tensor:
    """tiles: list of (1,C,scale*tile,scale*tile) outputs in grid order -> (1,C,scale*H,scale*W) with the padding removed."""
    H, W, ny, nx = meta
    C = tiles[0].shape[1]
    t = scale * tile
    full = tiles[0].new_empty((1, C, ny * t, nx * t))
    for i, x in enumerate(tiles):
        ty, tx = divmod(i, nx)
        full[:, :, ty * t:(ty + 1) * t, tx * t:(tx + 1) * t] = x
    return full[:, :, :scale * H, :scale * W]
