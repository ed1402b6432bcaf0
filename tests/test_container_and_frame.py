"""Host-side frame plumbing: Bayer packing, RCNB container, container merge.  The GPU round trip of a
ragged frame through compress_frame / decompress_frame is in test_gpu_parity.py."""
import pytest
import torch

from realcamnet_b200 import container, frame


def test_bayer_pack_is_inverse_pixel_shuffle():
    g = torch.Generator().manual_seed(3)
    mosaic = torch.randint(64, 1023, (12, 20), generator=g).to(torch.int16)
    p = frame.pack_bayer(mosaic, black_level=64, white_level=1023)
    assert tuple(p.shape) == (4, 6, 10) and float(p.min()) >= 0 and float(p.max()) <= 1
    assert torch.equal(torch.nn.functional.pixel_shuffle(p[None], 2)[0, 0], (mosaic.float() - 64) / (1023 - 64))
    assert torch.equal(frame.unpack_bayer(p), (mosaic.float() - 64) / (1023 - 64))
    with pytest.raises(ValueError):
        frame.pack_bayer(torch.zeros(3, 4))


def test_container_roundtrip_and_errors():
    hdr = container.FrameHeader(model_id=7, H=270, W=480, tile=128, ny=3, nx=4, n_tiles=3)
    recs = [container.TileStreams(11, (4, 4), b"\x01\x02\x03" * 100, b"zz"), container.TileStreams(0, (4, 4), b"", b"\x00" * 8),
            container.TileStreams(5, (2, 6), bytes(range(256)), b"q")]
    blob = container.pack(hdr, recs)
    h2, t2 = container.unpack(blob)
    assert h2 == hdr and set(t2) == {0, 5, 11}
    for r in recs:
        assert t2[r.index] == r
    bad = bytearray(blob)
    bad[40] ^= 1
    with pytest.raises(ValueError, match="CRC"):
        container.unpack(bytes(bad))
    with pytest.raises(ValueError):
        container.unpack(blob[:10])
    with pytest.raises(ValueError):
        container.pack(hdr, recs + [container.TileStreams(5, (1, 1), b"", b"")])     # count mismatch
    with pytest.raises(ValueError):
        container.pack(hdr._replace(n_tiles=2), [recs[0], recs[0]])                  # repeated index
    # two ranks' partial containers merge into the full one
    a = container.pack(hdr._replace(n_tiles=2), recs[:2])
    b = container.pack(hdr._replace(n_tiles=1), recs[2:])
    merged = frame.merge_containers([a, b])
    h3, t3 = container.unpack(merged)
    assert h3 == hdr and t3 == t2


class _StubCodec(torch.nn.Module):
    """Stands in for raw_compression_tcm_final in the HOST-logic test below: compress() serialises the tile and its coordinate map,
    decompress() returns the tile's first three channels upsampled x2 -- enough to check tile order, padding, coordinates, container
    and stitching of realcamnet_b200.frame without a GPU."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.coords = {}

    def compress(self, x):
        raw, cond, coord = x
        assert tuple(cond.shape) == (1, 4, 256, 256) and tuple(coord.shape) == (1, 2) + tuple(raw.shape[2:])
        y = raw.contiguous().numpy().tobytes()
        z = coord[0, :, [0, -1]][:, :, [0, -1]].contiguous().numpy().tobytes()      # the four corner coordinates
        return {"strings": [[y], [z]], "shape": torch.Size([raw.shape[2] // 64, raw.shape[3] // 64])}

    def decompress(self, strings, shape):
        T = int(shape[0]) * 64
        raw = torch.frombuffer(bytearray(strings[0][0]), dtype=torch.float32).reshape(1, 4, T, T)
        return {"x_hat": raw[:, :3].repeat_interleave(2, -1).repeat_interleave(2, -2)}


def test_frame_host_logic_with_stub_codec():
    g = torch.Generator().manual_seed(8)
    fr = torch.rand(4, 300, 600, generator=g)                      # 2 x 3 grid of 256-tiles, ragged right and bottom
    m = _StubCodec()
    blob = frame.compress_frame(m, fr, 256)
    parts = [frame.compress_frame(m, fr, 256, tile_indices=idx) for idx in ([0, 3], [4, 1], [5, 2])]   # three "ranks", any order
    assert frame.merge_containers(parts) == blob
    hdr, recs = container.unpack(blob)
    assert (hdr.H, hdr.W, hdr.tile, hdr.ny, hdr.nx, hdr.n_tiles) == (300, 600, 256, 2, 3, 6)
    # per-tile coordinate maps cover [-1, 1] over the REAL 300 x 600 frame (container v2): corners of tile 0 and tile 5; the
    # bottom-right corner of the padded grid (pixel 511, 767) lies outside the unit square
    c0 = torch.frombuffer(bytearray(recs[0].z), dtype=torch.float32).reshape(2, 2, 2)
    c5 = torch.frombuffer(bytearray(recs[5].z), dtype=torch.float32).reshape(2, 2, 2)
    assert float(c0[0, 0, 0]) == -1.0 and float(c0[1, 0, 0]) == -1.0
    assert abs(float(c5[0, 1, 1]) - (767 / 599 * 2 - 1)) < 1e-6 and abs(float(c5[1, 1, 1]) - (511 / 299 * 2 - 1)) < 1e-6
    with pytest.raises(ValueError):
        frame.decompress_frame(m, blob, model_id=7)                 # written by another model
    out = frame.decompress_frame(m, blob)
    assert tuple(out.shape) == (1, 3, 600, 1200)
    assert torch.equal(out[0], fr[:3].repeat_interleave(2, -1).repeat_interleave(2, -2))
    with pytest.raises(ValueError):
        frame.decompress_frame(m, parts[0])                         # a partial container cannot be decoded to a frame
    with pytest.raises(ValueError):
        frame.compress_frame(m, fr, 192)                            # tile side must be a multiple of 128 and >= 256
