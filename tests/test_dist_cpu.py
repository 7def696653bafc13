"""world_size-2/3 gloo tests (CPU) of the N>1 host logic: row ownership and the slab gather to rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rayz_b200 import _abi as abi
from rayz_b200.dist import SlabGather, shard_row_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, height, width, band, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = abi.load()
        rows = shard_row_indices(height, rank, world, band)
        # the C library and the python rule agree on how many rows this rank owns
        assert lib.rayz_cuda_shard_rows(height, rank, world, band) == len(rows)
        g = SlabGather(height, (width, 4), torch.float32, "cpu", band)
        assert g.my_rows() == len(rows)
        # synthetic slab: every pixel encodes its GLOBAL row and column, as a shard render would produce
        slab = torch.empty((len(rows), width, 4))
        for k, j in enumerate(rows):
            slab[k, :, 0] = j
            slab[k, :, 1] = torch.arange(width)
            slab[k, :, 2] = rank
            slab[k, :, 3] = 1
        for _ in range(2):  # reuse of the pre-allocated buffers across frames
            final = g.run(slab)
        if rank == 0:
            assert final.shape == (height, width, 4)
            assert torch.equal(final[:, 0, 0], torch.arange(height, dtype=torch.float32))
            assert torch.equal(final[3 % height, :, 1], torch.arange(width, dtype=torch.float32))
            owner = torch.tensor([(j // band) % world for j in range(height)], dtype=torch.float32)
            assert torch.equal(final[:, 0, 2], owner)
        else:
            assert final is None
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height,band", [(2, 37, 4), (3, 10, 1), (2, 3, 4)])
def test_slab_gather_gloo(world, height, band):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, 8, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_row_ownership_is_a_partition():
    for h in (1, 5, 675, 2160):
        for world in (1, 2, 4, 8):
            for band in (1, 4):
                seen = sorted(j for r in range(world) for j in shard_row_indices(h, r, world, band))
                assert seen == list(range(h))
