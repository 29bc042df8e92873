"""Multi-process host logic of the sharded paths on CPU: world_size 2 over gloo."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, height, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import flow2d_loader
    flow2d_loader.load()
    from cuda_flow2d_b200 import sharding
    mine = sharding.pairs_for_rank(n_pairs, rank, world)
    # every rank "processes" its pairs; gather the assignment and the slowest step time
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slabs = [None] * world
    dist.all_gather_object(slabs, sharding.slab_rows(height, rank, world))
    t = sharding.reduce_step_time(10.0 + 5.0 * rank, dist)
    dist.barrier()
    if rank == 0:
        out.put((gathered, slabs, t))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs,height", [(256, 8192), (7, 389)])
def test_two_rank_sharding(n_pairs, height):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, height, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, slabs, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for part in gathered for i in part)
    assert flat == list(range(n_pairs))                       # disjoint and covering
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1      # balanced
    assert slabs[0][0] == 0 and slabs[-1][1] == height and slabs[0][1] == slabs[1][0]
    assert abs((slabs[0][1] - slabs[0][0]) - (slabs[1][1] - slabs[1][0])) <= 1
    assert t == 15.0                                          # max over ranks


def test_sharding_single_process(pkg):
    from cuda_flow2d_b200 import sharding
    for world in (1, 2, 4, 8):
        parts = [sharding.pairs_for_rank(256, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(256))
        rows = [sharding.slab_rows(8192, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 8192
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.pairs_for_rank(4, 2, 2)
    assert sharding.reduce_step_time(3.5) == 3.5
