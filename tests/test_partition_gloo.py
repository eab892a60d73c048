"""Multi-rank frame assembly on CPU (gloo, world_size 2 and 3): each rank contributes only the
pixels of its tiles (zeros elsewhere, exactly what vimg_render writes with n_ranks > 1) and a
reduce(SUM) to rank 0 must rebuild the frame bit for bit.  The GPU version of the same property
is tests/test_gpu_parity.py::test_tile_partition_sums_to_the_full_frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import golden_ref, load_product


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vimg = load_product()
    full = golden_ref("cornell_small")["mis_64"]
    h, w, _ = full.shape
    owner = vimg.tile_owner(w, h, world)
    part = torch.from_numpy(np.where((owner == rank)[..., None], full, np.float32(0)).copy())
    dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(bool(np.array_equal(part.numpy(), full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_reduce_assembles_the_frame(world):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_tile_owner_partitions_the_frame(vimg):
    for (w, h, n) in [(160, 160, 2), (1800, 800, 8), (1366, 768, 3), (13, 9, 4)]:
        owner = vimg.tile_owner(w, h, n)
        assert owner.shape == (h, w) and owner.min() == 0 and owner.max() == min(n, ((w + 7) // 8) * ((h + 7) // 8)) - 1
        # the bottom-left 8x8 tile is tile 0; the tile above it is tile 1 (x-major, y fastest)
        assert owner[h - 1, 0] == 0
        if h > 8 and n > 1:
            assert owner[h - 9, 0] == 1


def _spp_worker(rank, world, port, q):
    """sample-range partition (VIMG_PARTITION_SPP): every rank owns the ranges rank, rank + world, ...; each rank's
    frame is (sum of its ranges' samples) / spp, and the reduce(SUM) is the whole estimate."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vimg = load_product()
    spp, streams = 37, 8
    rng = np.random.default_rng(7)
    samples = rng.random((spp, 6, 5, 3)).astype(np.float64)  # per-sample radiance of a tiny frame, same on every rank
    mine = vimg.spp_ranges(spp, streams, rank, world)
    part = torch.from_numpy(sum(samples[b:e].sum(0) for b, e in mine) / spp)
    counts = torch.tensor([sum(e - b for b, e in mine)], dtype=torch.int64)
    dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(counts, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(bool(np.allclose(part.numpy(), samples.mean(0), rtol=1e-12)) and int(counts[0]) == spp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_reduce_assembles_the_sample_range_partition(world):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_spp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_spp_ranges_cover_every_sample_once(vimg):
    for spp, k, n in [(512, 64, 8), (7, 3, 2), (33, 32, 4), (5, 64, 5), (1024, 16, 3)]:
        seen = np.zeros(spp, int)
        for r in range(n):
            for b, e in vimg.spp_ranges(spp, k, r, n):
                assert e > b
                seen[b:e] += 1
        assert (seen == 1).all()
    with pytest.raises(vimg.VimgError):
        vimg.spp_ranges(64, 2, 0, 3)
