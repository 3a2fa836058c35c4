"""CPU-only, world_size 2 over gloo: the N>1 plumbing of bench.py (channel sharding, max-over-ranks timing) without GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_channels(total, world, rank)
    mine = torch.zeros(total, dtype=torch.int64)
    mine[first:first + count] = 1
    dist.all_reduce(mine)
    elapsed = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    seeds = torch.tensor([sharding.channel_seed(7, first)], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, seeds)
    if rank == 0:
        out.put((mine.tolist(), float(elapsed.item()), [int(g.item()) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_every_channel_once():
    world, total = 2, 1025
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    cover, elapsed, seeds = out.get()
    assert cover == [1] * total
    assert elapsed == 11.0  # max over ranks
    assert seeds == [sharding.channel_seed(7, 0), sharding.channel_seed(7, 513)]


def test_shard_sizes():
    for total in (1, 7, 1024, 32768, 1025):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_channels(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
