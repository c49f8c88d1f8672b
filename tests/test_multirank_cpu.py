"""CPU tests of the N>1 path: the screen-tile partition and the packed-tile layout (host code
of the core: rt_tile_owners, rt_packed_floats) and the tile assembly protocol of
rt_render_multi -- pack own tiles, send to the root, scatter -- exercised with world_size-2
gloo process groups.  (The device halves, rt_render_tiles_packed / rt_unpack_tiles, are
checked on the GPU by tests/test_gpu_render.py::test_packed_tiles_gather_and_scatter.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("width,height,world,tile", [(3840, 2160, 8, 0), (100, 60, 3, 16), (33, 65, 2, 32), (7, 5, 4, 8)])
def test_partition_covers_every_tile_once_and_is_balanced(capi, width, height, world, tile):
    owners, ts = capi.tile_owners(width, height, world, tile)
    assert owners.shape == ((height + ts - 1) // ts, (width + ts - 1) // ts)
    assert owners.max() < world
    counts = np.bincount(owners.ravel(), minlength=world)
    assert counts.sum() == owners.size
    if owners.size >= 4 * world:
        assert counts.max() - counts.min() <= max(2, owners.size // (4 * world))
    # lattice interleave (tx + m ty) mod world: horizontally adjacent tiles belong to different ranks, and a rank's
    # own tiles keep their distance (for 8 ranks no two of them touch, not even corner to corner)
    if world > 1 and owners.shape[1] > 1:
        assert (owners[:, 1:] != owners[:, :-1]).all()
    if world == 8:
        assert (owners[1:, :] != owners[:-1, :]).all()
        assert (owners[1:, 1:] != owners[:-1, :-1]).all() and (owners[1:, :-1] != owners[:-1, 1:]).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, width, height, tile, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rayito_b200 import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owners, ts = capi.tile_owners(width, height, world, tile)
    # synthetic "render": every rank fills only the pixels of its own tiles with a
    # position-dependent value, exactly like rt_render_device leaves other pixels 0
    yy, xx = np.mgrid[0:height, 0:width]
    truth = np.stack([xx * 1.0, yy * 2.0, xx * yy * 0.5 + 1.0], axis=-1).astype(np.float32)
    mine = owners[yy // ts, xx // ts] == rank
    image = np.where(mine[..., None], truth, 0.0).astype(np.float32)
    # the path's one collective (rt_render_multi): every rank packs its tiles -- ascending tile
    # index, pixel (r, q) of the k-th tile at ((k * ts + r) * ts + q) * 3 -- and sends the packed
    # buffer to the root, which scatters each rank's tiles into the frame
    tiles_y, tiles_x = owners.shape

    def pack(r):
        ids = np.flatnonzero(owners.ravel() == r)
        buf = np.full((len(ids), ts, ts, 3), np.nan, np.float32)     # off-image pixels are never read
        for k, tid in enumerate(ids):
            y0, x0 = (tid // tiles_x) * ts, (tid % tiles_x) * ts
            h, w = min(ts, height - y0), min(ts, width - x0)
            buf[k, :h, :w] = image[y0:y0 + h, x0:x0 + w]
        return ids, buf

    ids, mine_packed = pack(rank)
    assert mine_packed.size == capi.packed_floats(width, height, world, rank, tile)
    frame = image.copy()
    if rank == 0:
        for r in range(1, world):
            n = capi.packed_floats(width, height, world, r, tile)
            got = torch.empty(n, dtype=torch.float32)
            if n:
                dist.recv(got, src=r)
            rids = np.flatnonzero(owners.ravel() == r)
            buf = got.numpy().reshape(len(rids), ts, ts, 3)
            for k, tid in enumerate(rids):
                y0, x0 = (tid // tiles_x) * ts, (tid % tiles_x) * ts
                h, w = min(ts, height - y0), min(ts, width - x0)
                frame[y0:y0 + h, x0:x0 + w] = buf[k, :h, :w]
    elif mine_packed.size:
        dist.send(torch.from_numpy(mine_packed.reshape(-1).copy()), dst=0)
    t = torch.from_numpy(frame)
    rays = torch.tensor([float(mine.sum())], dtype=torch.float64)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)       # bench.py's whole-job ray count
    elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)    # bench.py's max-over-ranks timing
    if rank == 0:
        np.save(out_path, t.numpy())
        assert rays.item() == width * height
        assert elapsed.item() == float(world)
    dist.destroy_process_group()


@pytest.mark.parametrize("width,height,tile", [(100, 60, 16), (64, 64, 32), (33, 65, 0)])
def test_tile_assembly_world2_gloo(tmp_path, width, height, tile):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "assembled.npy")
    mp.spawn(_worker, args=(2, port, width, height, tile, out), nprocs=2, join=True)
    got = np.load(out)
    yy, xx = np.mgrid[0:height, 0:width]
    truth = np.stack([xx * 1.0, yy * 2.0, xx * yy * 0.5 + 1.0], axis=-1).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), truth.view(np.uint32))
