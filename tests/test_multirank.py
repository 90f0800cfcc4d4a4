"""Host-side logic of the N>1 path with world_size 2 on CPU (gloo): contiguous batch sharding, the single weight-arena
broadcast, max-over-ranks timing, and the property the sharding relies on -- images are independent units, so running
the shards separately and concatenating equals running the whole batch (checked with the CPU oracle)."""
import os
import socket

import numpy as np
import pytest

from tengine_b200 import abi, sharding, workloads


def test_shard_ranges_partition_the_batch():
    for n in (1, 2, 7, 64, 128, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = sharding.shard_range(n, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(n)), (n, world)
            sizes = [sharding.shard_range(n, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle

        # (1) the one-time weight broadcast: rank 0 holds the packed arena, the others an empty one of the same size
        rng = np.random.default_rng(0)
        arena0 = rng.integers(0, 256, 4096 + 13).astype(np.uint8)
        arena = torch.from_numpy(arena0.copy() if rank == 0 else np.zeros_like(arena0))
        sharding.broadcast_arena(arena, src=0)
        ok_arena = bool(np.array_equal(arena.numpy(), arena0))
        # (2) every rank runs ITS slice of the global batch (no collective in the steady state)
        n_global = 5
        gfull, b = workloads.tiny_net(abi.DT_INT8, batch=n_global, seed=3)
        x = b.random_input(11)
        start, count = sharding.shard_range(n_global, world, rank)
        gshard, _ = workloads.tiny_net(abi.DT_INT8, batch=count, seed=3)
        y = Oracle().run(gshard, [x[start:start + count]])
        out = y[gshard.outputs[0]]
        # (3) timing reduction
        mx = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
        q.put((rank, ok_arena, start, out, mx))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharding_broadcast_and_timing(oracle):
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "weight arena differs after the broadcast"
    assert res[0][4] == [11.0, 5.0] and res[1][4] == [11.0, 5.0]
    # shards merged in rank order == the whole batch in one run
    gfull, b = workloads.tiny_net(abi.DT_INT8, batch=5, seed=3)
    x = b.random_input(11)
    want = oracle.run(gfull, [x])[gfull.outputs[0]]
    got = sharding.merge_shards([r[3] for r in res])
    assert np.array_equal(got, want)
