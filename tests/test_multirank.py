"""Host-side logic of the N > 1 path on CPU.  The product shards a batch over several GPUs inside ONE process (a tb200 context
over a GPU group, tests/test_gpu_multi.py); what can be checked without a GPU is (a) the library's sharding rule
(tb200_shard_range: contiguous slices of dim 0, sizes differing by at most one), (b) the property the sharding relies on --
images are independent units, so computing the slices separately and laying them side by side equals computing the whole batch --
with two processes (world_size 2, gloo) each taking the slice the library assigns to it and the CPU oracle as the checker, and
(c) the rank protocol of bench.py under torchrun: rank 0 drives the GPUs, every other rank only meets it at the barriers."""
import os
import socket

import numpy as np

from tengine_b200 import abi, workloads


def test_shard_ranges_partition_the_batch():
    from tengine_b200 import runtime as rt

    for n in (1, 2, 7, 64, 128, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = rt.shard_range(n, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(n)), (n, world)
            sizes = [rt.shard_range(n, world, r)[1] for r in range(world)]
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
        from tengine_b200 import runtime as rt

        # every rank computes ITS slice of the global batch (no collective on the data path) ...
        n_global = 5
        gfull, b = workloads.tiny_net(abi.DT_INT8, batch=n_global, seed=3)
        x = b.random_input(11)
        start, count = rt.shard_range(n_global, world, rank)
        gshard, _ = workloads.tiny_net(abi.DT_INT8, batch=count, seed=3)
        out = Oracle().run(gshard, [x[start:start + count]])[gshard.outputs[0]]
        # ... and the slices land in one buffer at the offsets the rule gives (all_gather stands in for the host buffer here)
        gathered = [None] * world
        dist.all_gather_object(gathered, (start, out))
        q.put((rank, gathered))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_slices_equal_the_whole_batch(oracle):
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
    gfull, b = workloads.tiny_net(abi.DT_INT8, batch=5, seed=3)
    x = b.random_input(11)
    want = oracle.run(gfull, [x])[gfull.outputs[0]]
    for rank, gathered in res:
        got = np.zeros_like(want)
        for start, out in gathered:
            got[start:start + out.shape[0]] = out
        assert np.array_equal(got, want), rank


def _bench_rank(rank, world, port, q):
    """bench.py's own rank protocol (the code the driver's torchrun launch executes), without a GPU: the leader announces each of
    its barriers, the followers meet it there and leave when told -- whatever the number of barriers."""
    import sys

    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    if rank == 0:
        lead = bench.RankLead(world, [0])
        try:
            for _ in range(5):
                lead.barrier()
        finally:
            lead.finish()
        q.put((rank, 5))
    else:
        q.put((rank, bench.rank_follow(world)))


def test_bench_rank_protocol_world_size_3_gloo():
    import torch.multiprocessing as mp

    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bench_rank, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == {0: 5, 1: 5, 2: 5}  # every follower met the leader at each of its barriers and then left
