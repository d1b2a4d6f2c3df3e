"""N>1 host logic on CPU: world_size-2 gloo run of the frame sharding + size gather (no data-path collective exists)."""
import os
import socket

import pytest

from jpgenc_b200.sharding import frames_for_rank, offsets_from_sizes, owner_of


@pytest.mark.parametrize("n,world", [(1024, 8), (1024, 3), (5, 8), (0, 2), (7, 7), (10, 4)])
def test_partition_is_exact_and_balanced(n, world):
    seen = []
    sizes = []
    for r in range(world):
        fr = frames_for_rank(n, r, world)
        seen.extend(fr)
        sizes.append(len(fr))
        for f in fr:
            assert owner_of(f, n, world) == r
    assert seen == list(range(n))
    assert max(sizes) - min(sizes) <= 1


def test_offsets():
    assert offsets_from_sizes([3, 0, 5]) == [(0, 3), (3, 0), (3, 5)]


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from jpgenc_b200.sharding import frames_for_rank, gather_frame_sizes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = frames_for_rank(n_frames, rank, world)
    sizes = gather_frame_sizes([1000 + 7 * f for f in mine], n_frames)     # stand-in for encoded sizes
    q.put((rank, sizes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames = 9
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    expect = [1000 + 7 * f for f in range(n_frames)]
    assert got[0] == expect and got[1] == expect
