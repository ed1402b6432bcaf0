"""Frame-level pipeline around the tile codec (BASELINE config 4):

  Bayer mosaic (2H x 2W sensor samples) -> packed 4-channel frame (4,H,W) -> zero-padded tiles + per-tile coordinate maps
  -> raw_compression_tcm_final.compress per tile -> RCNB container;   and back: container -> per-tile decompress -> stitched sRGB.

The reference has no code for these steps (its smoke tests feed random tensors of the tile shapes, models/LiteISP.py:2670-2672);
the conventions follow its architecture figure ("unpixel shuffle", coordinate map) and the padding policy of
``pad_to_multiple_of_16`` (models/LiteISP.py:84-105).  Everything here is host-side indexing; the arithmetic stays in the tile
codec's CUDA kernels.  Multi-GPU: tiles are independent, tile t belongs to rank t mod G (realcamnet_b200.dist).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np
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


# ----------------------------------------------------------------------------- batched, pipelined tile compression
_pool = None


def _coder_pool() -> ThreadPoolExecutor:
    """Host range-coder workers.  The coder is a C call through ctypes (the GIL is released while it runs), one independent
    stream per tile: the streams of a batch are encoded in parallel, and while the GPU already runs the next batch."""
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) - 1)), thread_name_prefix="rcn-coder")
    return _pool


MAX_BATCH = int(os.environ.get("RCN_FRAME_MAX_BATCH", "8"))     # tiles per graph replay (a tile's streams do not depend on it)


def batch_size_for(n_tiles: int, max_batch: int = 0) -> int:
    """Largest divisor of n_tiles that is <= max_batch (default MAX_BATCH): equal batches, one captured graph per rank."""
    max_batch = max_batch or MAX_BATCH
    for b in range(min(max_batch, n_tiles), 0, -1):
        if n_tiles % b == 0:
            return b
    return 1


def _encode_tile(model, pending, n: int, B: int) -> Tuple[bytes, bytes]:
    """Host side of one tile of a batch: its y stream from the GPU-prepared coder input (slice-major, then this tile's (c,h,w)
    block of every slice -- exactly the order compress() of the single tile produces) and its z stream."""
    from .entropy_models import rans_encode_packed

    done, host, _keep = pending
    done.synchronize()
    h_packed, h_raw, h_flags, h_zsym = host
    S = model.num_slices
    pk = np.ascontiguousarray(h_packed.numpy().reshape(S, B, -1)[:, n]).reshape(-1)
    rw = np.ascontiguousarray(h_raw.numpy().reshape(S, B, -1)[:, n]).reshape(-1)
    fl = np.ascontiguousarray(h_flags.numpy().reshape(S, B, -1)[:, n]).reshape(-1)
    y = rans_encode_packed(pk, rw, fl)
    z = model.entropy_bottleneck.compress_symbols(h_zsym[n:n + 1])[0]
    return y, z


@torch.no_grad()
def compress_tiles(model, tiles: torch.Tensor, cond: torch.Tensor, coords: torch.Tensor, max_batch: int = 0):
    """tiles (n,4,T,T), cond (1,4,h,w), coords (n,2,T,T) on the model's device -> [(y_bytes, z_bytes, (zh, zw))] * n.

    Tiles are independent images, so they are pushed through the compress stage in equal batches of up to `max_batch` (one CUDA
    graph replay per batch: at 512-tiles the ~600 launches of a single tile are launch-latency bound), and the host range coder
    of batch k runs on worker threads while the GPU computes batch k+1 (double-buffered pinned staging).  Every tile's streams are
    byte-identical to model.compress() of that tile alone (tests/test_gpu_parity.py::test_tile_pipeline_equals_per_tile_compress)."""
    from .tcm import StageRunner

    n = tiles.shape[0]
    if n == 0:
        return []
    if model.gaussian_conditional._offset.numel() == 0:
        raise RuntimeError("call update() before compressing")
    B = batch_size_for(n, max_batch)
    pool = _coder_pool()
    jobs, copy_done = [], None
    model.entropy_bottleneck.host_tables()       # build the host tables on this thread, not racily on the workers
    for k in range(n // B):
        sl = slice(k * B, (k + 1) * B)
        xs = [tiles[sl], cond.expand(B, -1, -1, -1), coords[sl]]
        if k >= 2:                               # the staging slot about to be refilled must have been consumed
            for f in jobs[(k - 2) * B:(k - 1) * B]:
                f.result()
        if copy_done is not None:                # the stage's outputs live in the graph pool: finish the previous copy first
            torch.cuda.current_stream().wait_event(copy_done)
        E = StageRunner(model, ("compress",), xs, model._compress_stage, None).a()
        pending = model._begin_host_copy(E.coder.packed, E.coder.raw, E.coder.flags, E.z_sym, slot=k % 2)
        copy_done = pending[0]
        shape = (int(E.z.shape[1]), int(E.z.shape[2]))
        jobs.extend(pool.submit(_encode_tile, model, pending, i, B) for i in range(B))
    return [f.result() + (shape,) for f in jobs]


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
    if hasattr(model, "_compress_stage") and dev.type == "cuda":      # the codec models: batched + pipelined
        out = compress_tiles(model, tiles[todo].to(dev), cond, tiler.tiles_coords(meta, tile, todo, device=dev))
        recs = [container.TileStreams(t, shp, y, z) for t, (y, z, shp) in zip(todo, out)]
    else:                                                             # any object with the reference's compress() signature
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
