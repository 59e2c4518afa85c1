"""Frame-block partition and frame gathering with world_size 2 on CPU (gloo)."""
import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scopyon_b200.movie import assemble_channels, frame_block, gather_frames


def test_frame_blocks_cover_the_movie():
    for frames in (1, 7, 10000, 30):
        for world in (1, 2, 3, 8):
            blocks = [frame_block(frames, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == frames
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_frames, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = frame_block(num_frames, rank, world)
    # frame f is filled with the value f: the stack must come back in frame order
    local = torch.stack([torch.full((4, 3), float(f)) for f in range(first, last)]) if last > first \
        else torch.zeros((0, 4, 3))
    everywhere = gather_frames(local, num_frames)
    on_root = gather_frames(local, num_frames, dst=0)
    channels = assemble_channels(torch.full((4, 3), float(rank + 1)))
    ok = everywhere.shape == (num_frames, 4, 3) and all(
        bool((everywhere[f] == f).all()) for f in range(num_frames))
    ok = ok and ((on_root is None) == (rank != 0))
    if rank == 0:
        ok = ok and torch.equal(on_root, everywhere)
    ok = ok and channels.shape == (4, 3, world) and all(bool((channels[..., r] == r + 1).all()) for r in range(world))
    numpy.save(os.path.join(result_dir, "ok{}.npy".format(rank)), numpy.array(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [7, 2])
def test_gather_frames_gloo_world2(tmp_path, num_frames):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), num_frames, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(numpy.load(tmp_path / "ok{}.npy".format(r)))
