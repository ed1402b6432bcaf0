"""Frame-level pipeline around the tile codec (BASELINE config 4):

  Bayer mosaic (2H x 2W sensor samples) -> packed 4-channel frame (4,H,W) -> zero-padded tiles + per-tile coordinate maps
  -> raw_compression_tcm_final.compress per tile -> RCNB container;   and back: container -> per-tile decompress -> stitched sRGB.

The reference has no code for these steps (its smoke tests feed random tensors of the tile shapes, models/LiteISP.py:2670-2672);
the conventions follow its architecture figure ("unpixel shuffle", coordinate map) and the padding policy of
``pad_to_multiple_of_16`` (models/LiteISP.py:84-105).  Everything here is host-side indexing; the arithmetic stays in the tile
codec's CUDA kernels.  Multi-GPU: tiles are independent, tile t belongs to rank t mod G (realcamnet_b200.dist).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import container, tiler


def pack_bayer(mosaic: torch.Tensor, black_level: float = 0.0, white_level: float = 1.0) -> torch.Tensor:
    """(2H,2W) or (1,2H,2W) sensor mosaic (any real/integer dtype) -> (4,H,W) float32 in [0,1]: channel 2*i+j holds the
    photosites at (2y+i, 2x+j) -- the inverse of nn.PixelShuffle(2) ("unpixel shuffle")."""
    if mosaic.dim() == 3:
        mosaic = mosaic[0]
    H2, W2 = mosaic.shape
    if H2 % 2 or W2 % 2:
        raise ValueError("Bayer mosaic needs even dimensions")
    m = (mosaic.to(torch.float32) - black_level) / (white_level - black_level)
    return m.reshape(H2 // 2, 2, W2 // 2, 2).permute(1, 3, 0, 2).reshape(4, H2 // 2, W2 // 2).clamp_(0, 1).contiguous()


def unpack_bayer(packed: torch.Tensor) -> torch.Tensor:
    """(4,H,W) -> (2H,2W): inverse of pack_bayer (without the level scaling)."""
    _, H, W = packed.shape
    return packed.reshape(2, 2, H, W).permute(2, 0, 3, 1).reshape(2 * H, 2 * W).contiguous()


def frame_condition(frame: torch.Tensor, size: int = 256) -> torch.Tensor:
    """The global colour-condition input: the whole packed frame resized to size x size (bilinear), (1,4,size,size).
    Input preparation, done once per frame."""
    return torch.nn.functional.interpolate(frame[None] if frame.dim() == 3 else frame, size=(size, size), mode="bilinear",
                                           align_corners=False)


def compress_frame(model, frame: torch.Tensor, tile: int, cond: Optional[torch.Tensor] = None, model_id: int = 0,
                   tile_indices: Optional[List[int]] = None) -> bytes:
    """frame (4,H,W) packed Bayer in [0,1] -> RCNB container bytes.  tile must be a multiple of 128 and at least 256 (window /
    stride divisibility of the codec: the latent map at tile/16 must be larger than the 8x8 attention window, SURVEY.md
    section 6).  tile_indices restricts the work to a subset (one rank's share)."""
    if tile % 128 or tile < 256:
        raise ValueError("tile side must be a multiple of 128 and >= 256")
    dev = next(model.parameters()).device
    tiles, meta = tiler.split_frame(frame, tile)
    H, W, ny, nx = meta
    if cond is None:
        cond = frame_condition(frame)
    cond = cond.to(dev)
    todo = list(range(ny * nx)) if tile_indices is None else list(tile_indices)
    recs = []
    for t in todo:
        x = [tiles[t:t + 1].to(dev), cond, tiler.tile_coords(meta, tile, t, device=dev)]
        c = model.compress(x)
        recs.append(container.TileStreams(t, tuple(int(v) for v in c["shape"]), c["strings"][0][0], c["strings"][1][0]))
    return container.pack(container.FrameHeader(model_id, H, W, tile, ny, nx, len(recs)), recs)


def decompress_frame(model, blob: bytes, model_id: int = 0) -> torch.Tensor:
    """RCNB container -> (1,3,2H,2W) sRGB frame in [0,1] (padding removed).  Every tile of the grid must be present.
    The container is untrusted input (CRC-32 is not authentication): geometry fields are validated before anything is
    allocated from them."""
    hdr, tiles = container.unpack(blob)
    if hdr.model_id != model_id:
        raise ValueError(f"container was written by model {hdr.model_id}, this decoder is model {model_id}")
    if hdr.tile % 128 or hdr.tile < 256 or hdr.tile > 8192:
        raise ValueError(f"container tile side {hdr.tile} is not a valid codec tile")
    if (hdr.ny, hdr.nx) != tiler.tile_grid(hdr.H, hdr.W, hdr.tile) or hdr.H <= 0 or hdr.W <= 0:
        raise ValueError("container grid does not match its frame size")
    if len(tiles) != hdr.ny * hdr.nx:
        raise ValueError(f"container holds {len(tiles)} of {hdr.ny * hdr.nx} tiles")
    outs = []
    for t in range(hdr.ny * hdr.nx):
        rec = tiles[t]
        if tuple(rec.shape) != (hdr.tile // 64, hdr.tile // 64):
            raise ValueError(f"tile {t}: hyper-latent shape {rec.shape} does not belong to a {hdr.tile}-tile")
        outs.append(model.decompress([[rec.y], [rec.z]], rec.shape)["x_hat"])
    return tiler.stitch(outs, (hdr.H, hdr.W, hdr.ny, hdr.nx), hdr.tile, scale=2)


def merge_containers(blobs: List[bytes]) -> bytes:
    """Joins the partial containers produced by several ranks (disjoint tile subsets of one frame)."""
    hdr0, recs = None, []
    for b in blobs:
        hdr, tiles = container.unpack(b)
        if hdr0 is None:
            hdr0 = hdr
        elif hdr[:6] != hdr0[:6]:
            raise ValueError("containers describe different frames")
        recs.extend(tiles.values())
    recs.sort(key=lambda r: r.index)
    return container.pack(hdr0._replace(n_tiles=len(recs)), recs)
