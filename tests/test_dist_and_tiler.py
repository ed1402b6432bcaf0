"""CPU tests of the multi-GPU host logic: world_size-2 gloo processes exercise the tile scatter and the
variable-length bitstream gather; the tiler round-trips a ragged frame."""
import hashlib
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realcamnet_b200 import dist as rdist
from realcamnet_b200 import tiler


def test_tiler_roundtrip_ragged_frame():
    g = torch.Generator().manual_seed(0)
    frame = torch.rand(4, 2160 // 8, 3840 // 8, generator=g)       # 270 x 480: not a multiple of the tile
    tiles, meta = tiler.split_frame(frame, 64)
    assert meta == (270, 480, 5, 8) and tuple(tiles.shape) == (40, 4, 64, 64)      # BASELINE config 4 grid: 8 x 5 = 40 tiles
    assert float(tiles[39, :, 270 - 256:, :].abs().sum()) == 0.0                   # bottom padding is zero
    up = [t[None].repeat_interleave(2, -1).repeat_interleave(2, -2) for t in tiles]
    full = tiler.stitch(up, meta, 64, scale=2)
    assert tuple(full.shape) == (1, 4, 540, 960)
    assert torch.equal(full[0, :, ::2, ::2], frame)
    c0, c39 = tiler.tile_coords(meta, 64, 0), tiler.tile_coords(meta, 64, 39)
    # coordinates are normalised by the REAL frame (270 x 480): the last real pixel is +1, padded pixels fall outside [-1, 1],
    # and the frame centre is (0, 0) whatever the tile size
    assert c0[0, 0, 0, 0] == -1 and c0[0, 1, 0, 0] == -1
    assert abs(float(c39[0, 0, 0, 479 - 448]) - 1) < 1e-6 and abs(float(c39[0, 1, 269 - 256, 0]) - 1) < 1e-6
    assert float(c39[0, 0, -1, -1]) > 1 and float(c39[0, 1, -1, -1]) > 1
    big = tiler.tile_coords((270, 480, 1, 1), 512, 0)
    assert torch.allclose(big[0, :, :64, :64], c0[0], atol=1e-6)
    assert rdist.my_tiles(40, 3, 8) == [3, 11, 19, 27, 35]


def _fake_codec(tile, index):
    """Deterministic, variable-length stand-in for model.compress on one tile."""
    h = hashlib.sha256(tile.numpy().tobytes()).digest()
    return h * (1 + index % 3) + bytes([index])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        tiles = torch.rand(7, 4, 8, 8, generator=g)                   # 7 tiles over 2 ranks: 4 + 3
        mine = rdist.scatter_tiles(tiles if rank == 0 else None, 7, (4, 8, 8))
        assert torch.equal(mine, tiles[rdist.my_tiles(7, rank, world)])
        out = rdist.compress_frame_sharded(tiles if rank == 0 else None, 7, (4, 8, 8), _fake_codec)
        if rank == 0:
            want = [_fake_codec(tiles[i:i + 1], i) for i in range(7)]
            q.put(out == want)
        else:
            q.put(out is None)
    finally:
        dist.destroy_process_group()


def test_two_rank_scatter_and_bitstream_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)


def _fake_compress_tiles(model, tiles, cond, coords, max_batch=8):
    """Stand-in for frame.compress_tiles (the CUDA codec): streams are hashes of the tile, its coordinates and the condition."""
    out = []
    for i in range(tiles.shape[0]):
        h = hashlib.sha256(tiles[i].numpy().tobytes() + coords[i].numpy().tobytes() + cond.numpy().tobytes()).digest()
        out.append((h * (1 + h[0] % 3), h[:7], (tiles.shape[2] // 64, tiles.shape[3] // 64)))
    return out


def _frame_worker(rank, world, port, q):
    import realcamnet_b200.frame as rframe

    rframe.compress_tiles = _fake_compress_tiles
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fr = torch.rand(4, 300, 600, generator=torch.Generator().manual_seed(21)) if rank == 0 else None
        blob = rdist.compress_frame_distributed(None, fr, 300, 600, 128, torch.device("cpu"))
        q.put((rank, blob))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_frame_container_is_identical_for_one_and_two_ranks():
    """BASELINE config 4 host logic on CPU (gloo): scatter, condition broadcast, per-rank tile codec, variable-length gather and
    container packing give the same bytes whatever the number of ranks."""
    from realcamnet_b200 import container

    ctx = mp.get_context("spawn")
    blobs = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = 31500 + (os.getpid() % 2000) + world
        procs = [ctx.Process(target=_frame_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=180) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(res[r] is None for r in range(1, world))
        blobs[world] = res[0]
    assert blobs[1] == blobs[2]
    hdr, recs = container.unpack(blobs[1])
    assert (hdr.H, hdr.W, hdr.tile, hdr.ny, hdr.nx, hdr.n_tiles) == (300, 600, 128, 3, 5, 15) and len(recs) == 15
