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
