"""Host-side multi-GPU logic on CPU: round-robin batch sharding and the 2-rank (gloo) throughput reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bisinger_b200.shard import make_batches, reduce_throughput, shard_batches


def test_make_batches_ragged_and_empty():
    assert make_batches(0, 32) == []
    assert make_batches(70, 32) == [(0, 32), (32, 64), (64, 70)]
    with pytest.raises(ValueError):
        make_batches(10, 0)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shards_partition_cfg5(world):
    # cfg5: 4096 phrases in batches of 32 -> 128 batches, disjoint cover, balanced
    all_b = make_batches(4096, 32)
    got = []
    for r in range(world):
        mine = shard_batches(4096, 32, r, world)
        assert len(mine) == 128 // world
        got += mine
    assert sorted(got) == all_b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_batches(100, 8, rank, world)              # 13 batches, ragged tail
    audio = sum((b - a) * 10.0 for a, b in mine)           # 10 s phrases
    total, tmax = reduce_throughput(audio, 1.0 + rank)
    dist.barrier()
    q.put((rank, len(mine), total, tmax))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [7, 6]
    for r in res:
        assert r[2] == 1000.0 and r[3] == 2.0               # sum of audio seconds, max of elapsed
