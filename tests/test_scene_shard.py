"""CPU / gloo, world_size 2: the N>1 host logic (scene sharding, detection gather, max-over-ranks time)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lib.model.scene_shard import gather_detections, max_over_ranks, shard_scenes


def test_shard_is_a_partition_and_balanced():
    costs = [96 * 48 * 96, 88 * 44 * 88, 208 * 48 * 160, 85 * 43 * 85] * 13 + [1000]
    for world in (1, 2, 4, 8):
        parts = [shard_scenes(costs, r, world) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [10, 40, 20, 30, 50]
    mine = shard_scenes(costs, rank, world)
    local = []
    for i in mine:
        k = i % 3  # scene 0 and 3 have no detections
        local.append((i, torch.full((k, 6), float(i)), torch.full((k,), i, dtype=torch.long), torch.full((k,), 0.5 + 0.01 * i)))
    allres = gather_detections(local)
    t = max_over_ranks(10.0 + rank)
    q.put((rank, mine, {k: (v[0].tolist(), v[1].tolist()) for k, v in allres.items()}, t))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert sorted(res[0][1] + res[1][1]) == [0, 1, 2, 3, 4]
    assert res[0][3] == res[1][3] == 11.0
    for _, _, allres, _ in res:  # every rank sees every scene's detections
        assert len(allres[4][0]) == 1 and allres[4][1] == [4]
        assert len(allres[2][0]) == 2 and len(allres[1][0]) == 1
        assert allres[0][0] == [] and allres[3][0] == []
