"""The N>1 host path on CPU: world_size-2 gloo process group, streams sharded by size, results gathered in the
original order.  The per-rank worker here is the oracle (the CUDA worker needs a GPU); the sharding, gather and
max-over-ranks timing logic is the code bench.py and deft4j_b200.sharding run on the GPU box."""
import os
import sys

import pytest

from conftest import ROOT


def test_shard_streams_is_a_balanced_partition():
    from deft4j_b200.sharding import shard_streams
    import random
    rnd = random.Random(3)
    for world in (1, 2, 4, 8):
        sizes = [rnd.randint(1, 1 << 20) for _ in range(rnd.randint(0, 200))]
        shards = shard_streams(sizes, world)
        assert sorted(i for s in shards for i in s) == list(range(len(sizes)))
        if len(sizes) >= 8 * world:
            loads = [sum(sizes[i] for i in s) for s in shards]
            assert max(loads) - min(loads) <= max(sizes)
    assert shard_streams([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]
    assert shard_streams([], 4) == [[], [], [], []]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import workloads as W
    import oracle_lib
    from deft4j_b200.sharding import optimise_sharded, rank_world
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert rank_world() == (rank, world)

    def oracle_worker(bufs, merge):
        out = []
        for raw in bufs:
            s = oracle_lib.OracleDeflateStream()
            assert s.parse(raw)
            saved = s.optimise(merge)
            out.append({"status": 0, "saved_bits": saved, "out": s.asBytes(), "rank": rank})
        return out

    streams = W.c4_streams(7, seed=21) + list(W.handmade_streams().values())[:3]
    res = optimise_sharded(streams, True, worker=oracle_worker)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction bench.py uses
    q.put((rank, [(r["saved_bits"], r["out"], r["rank"]) for r in res], t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_optimise():
    import torch.multiprocessing as mp
    import socket
    import workloads as W
    import oracle_lib
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = dict()
    for _ in range(2):
        rank, res, tmax = q.get(timeout=300)
        got[rank] = res
        assert tmax == 2.0
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0] == got[1]  # every rank sees the full, identically ordered result list
    streams = W.c4_streams(7, seed=21) + list(W.handmade_streams().values())[:3]
    assert {r[2] for r in got[0]} == {0, 1}  # both ranks did work
    for raw, (saved, out, _) in zip(streams, got[0]):
        o = oracle_lib.OracleDeflateStream(); assert o.parse(raw)
        assert o.optimise(True) == saved and o.asBytes() == out
