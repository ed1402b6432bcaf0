"""Bitstream container for tiled RAW frames.

The reference returns bare Python lists -- ``{"strings": [[y_bytes], [z_bytes]], "shape": (h, w)}`` per call
(models/raw2bit.py:1960) -- which cannot be stored or sent.  This module defines the length-prefixed wire format the frame
pipeline uses; it is host-side byte packing only (little-endian, no alignment padding):

  header   "RCNB" | u16 version=2 | u16 model_id | u32 frame_H | u32 frame_W | u16 tile | u16 grid_ny | u16 grid_nx | u32 n_tiles
  per tile u32 index | u16 z_h | u16 z_w | u32 len_y | u32 len_z | y bytes | z bytes
  trailer  u32 CRC-32 of everything before it

frame_H / frame_W are the packed-Bayer dimensions before padding; tiles appear in any order (index = ty * grid_nx + tx).
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, NamedTuple, Tuple

MAGIC = b"RCNB"
VERSION = 2       # v2: tile coordinate maps are normalised by the unpadded frame size (tiler.tile_coords)
_HEADER = struct.Struct("<4sHHIIHHHI")
_TILE = struct.Struct("<IHHII")


class TileStreams(NamedTuple):
    index: int
    shape: Tuple[int, int]      # the "shape" entry of compress(): hyper-latent (h, w)
    y: bytes
    z: bytes


class FrameHeader(NamedTuple):
    model_id: int
    H: int
    W: int
    tile: int
    ny: int
    nx: int
    n_tiles: int


def pack(header: FrameHeader, tiles: List[TileStreams]) -> bytes:
    if header.n_tiles != len(tiles):
        raise ValueError(f"header says {header.n_tiles} tiles, got {len(tiles)}")
    seen = set()
    parts = [_HEADER.pack(MAGIC, VERSION, header.model_id, header.H, header.W, header.tile, header.ny, header.nx, header.n_tiles)]
    for t in tiles:
        if not 0 <= t.index < header.ny * header.nx or t.index in seen:
            raise ValueError(f"tile index {t.index} out of range or repeated")
        seen.add(t.index)
        parts.append(_TILE.pack(t.index, int(t.shape[0]), int(t.shape[1]), len(t.y), len(t.z)))
        parts.append(bytes(t.y))
        parts.append(bytes(t.z))
    body = b"".join(parts)
    return body + struct.pack("<I", zlib.crc32(body) & 0xFFFFFFFF)


def unpack(blob: bytes) -> Tuple[FrameHeader, Dict[int, TileStreams]]:
    if len(blob) < _HEADER.size + 4:
        raise ValueError("container truncated")
    body, (crc,) = blob[:-4], struct.unpack("<I", blob[-4:])
    if zlib.crc32(body) & 0xFFFFFFFF != crc:
        raise ValueError("container CRC mismatch")
    magic, version, model_id, H, W, tile, ny, nx, n = _HEADER.unpack_from(body, 0)
    if magic != MAGIC or version != VERSION:
        raise ValueError(f"not an RCNB v{VERSION} container")
    off = _HEADER.size
    tiles: Dict[int, TileStreams] = {}
    for _ in range(n):
        if off + _TILE.size > len(body):
            raise ValueError("container truncated inside a tile record")
        index, zh, zw, ly, lz = _TILE.unpack_from(body, off)
        off += _TILE.size
        if off + ly + lz > len(body) or index in tiles or index >= max(ny * nx, 1):
            raise ValueError("container tile record is inconsistent")
        tiles[index] = TileStreams(index, (zh, zw), body[off:off + ly], body[off + ly:off + ly + lz])
        off += ly + lz
    if off != len(body):
        raise ValueError("trailing bytes in container")
    return FrameHeader(model_id, H, W, tile, ny, nx, n), tiles
