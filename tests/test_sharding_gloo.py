"""N > 1 host logic on CPU: contiguous frame blocks and the in-place all-gather of velocity rows,
world_size 2 and 3 over gloo (127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radar_slam_b200.sharding import frame_block, max_block, VelocityGather


def test_frame_block_partition():
    for world in (1, 2, 3, 8):
        for total in (0, 1, 7, 8, 1000, 65536, 65537):
            blocks = [frame_block(r, world, total) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == (max_block(world, total) if total else 0)
    with pytest.raises(ValueError):
        frame_block(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = VelocityGather(total, device="cpu")
        lo, hi = frame_block(rank, world, total)
        # what the velocity solve would write for frames [lo, hi): row k = (frame index, rank, ...)
        slot = g.slot()
        slot[: hi - lo, 0] = torch.arange(lo, hi, dtype=torch.float64)
        slot[: hi - lo, 1] = float(rank)
        slot[: hi - lo, 6] = 1.0
        g.gather()
        full = g.assemble()
        np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 10), (2, 7), (3, 8)])
def test_velocity_all_gather_gloo(tmp_path, world, total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for o in outs:
        assert o.shape == (total, 8)
        assert np.array_equal(o, outs[0])                 # every rank holds the same gathered sequence
    assert np.array_equal(outs[0][:, 0], np.arange(total))    # frame order restored
    for r in range(world):
        lo, hi = frame_block(r, world, total)
        assert np.all(outs[0][lo:hi, 1] == r) and np.all(outs[0][lo:hi, 6] == 1.0)
