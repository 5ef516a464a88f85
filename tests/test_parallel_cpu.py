"""world_size-2 gloo tests of the frame-sharding host logic, run on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stable_diffusion_videos_b200.parallel import broadcast_state_dict, frame_block, gather_frames


def test_frame_block_covers_everything_once():
    for n in (0, 1, 7, 60, 120, 360):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = frame_block(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    assert frame_block(60, 8, 7) == (56, 60)  # cfg-4: 8,8,...,4 -> 93.75 % balance


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weights: rank 0 owns the real values, others garbage -> broadcast makes them equal
        sd = {"b": torch.full((3,), float(rank)), "a": torch.arange(4.0).reshape(2, 2) + rank}
        sd = broadcast_state_dict(sd, src=0)
        ok_w = bool((sd["b"] == 0).all() and torch.equal(sd["a"].float(), torch.arange(4.0).reshape(2, 2)))
        lo, hi = frame_block(n_total, world, rank)
        # "render": frame i is filled with value i
        local = torch.stack([torch.full((4, 4, 3), i, dtype=torch.uint8) for i in range(lo, hi)]) if hi > lo else \
            torch.zeros((0, 4, 4, 3), dtype=torch.uint8)
        out = gather_frames(local, n_total, dst=0)
        if rank == 0:
            ok = out.shape == (n_total, 4, 4, 3) and all(int(out[i, 0, 0, 0]) == i for i in range(n_total))
            q.put(bool(ok and ok_w))
        else:
            q.put(bool(out is None and ok_w))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8, 1])
def test_gather_frames_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(res)
