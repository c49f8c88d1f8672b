"""CPU tests of the N>1 path: the screen-tile partition (host code of the core) and
the tile assembly collective, exercised with world_size-2 gloo process groups."""
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
    # diagonal interleave: horizontally and vertically adjacent tiles belong to different ranks
    if world > 1 and owners.shape[1] > 1:
        assert (owners[:, 1:] != owners[:, :-1]).all()


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
    t = torch.from_numpy(image.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)      # the path's one collective: tile assembly
    rays = torch.tensor([float(mine.sum())], dtype=torch.float64)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)       # bench.py's whole-job ray count
    elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)    # bench.py's max-over-ranks timing
    if rank == 0:
        np.save(out_path, t.numpy())
        assert rays.item() == width * height
        assert elapsed.item() == float(world)
    dist.destroy_process_group()


@pytest.mark.parametrize("width,height,tile", [(100, 60, 16), (64, 64, 32)])
def test_tile_assembly_world2_gloo(tmp_path, width, height, tile):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / "assembled.npy")
    mp.spawn(_worker, args=(2, port, width, height, tile, out), nprocs=2, join=True)
    got = np.load(out)
    yy, xx = np.mgrid[0:height, 0:width]
    truth = np.stack([xx * 1.0, yy * 2.0, xx * yy * 0.5 + 1.0], axis=-1).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), truth.view(np.uint32))
