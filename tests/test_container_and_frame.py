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
