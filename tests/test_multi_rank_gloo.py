"""world_size-2 run of the multi-GPU plumbing on CPU (gloo): shards partition the work exactly,
timings reduce with MAX, counts with SUM, and the weak-scaling throughput formula is the one
bench.py prints.  The data path itself has no collective to test."""
import os
import socket

import torch
import torch.multiprocessing as mp

from sast_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, l = parallel.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(17, r, w)
    streams = parallel.assign_streams(5, w)[r]
    parallel.barrier()
    t_max, = parallel.reduce_scalars([0.010 * (r + 1)], "max")
    n_sum, = parallel.reduce_scalars([hi - lo], "sum")
    q.put((rank, lo, hi, streams, t_max, n_sum, parallel.throughput(8, 20, t_max, w)))
    torch.distributed.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, t0, n0, f0), (r1, lo1, hi1, s1, t1, n1, f1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 9, 9, 17)                 # contiguous, remainder to the first rank
    assert s0 == [0, 2, 4] and s1 == [1, 3]                       # streams never move between ranks
    assert t0 == t1 == 0.020                                      # slowest rank defines the step
    assert n0 == n1 == 17
    assert f0 == f1 == 8 * 2 * 20 / 0.020


def test_sharding_properties():
    for n in (0, 1, 7, 8, 64, 1000):
        for world in (1, 2, 3, 4, 8):
            cuts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
            owned = sorted(s for lst in parallel.assign_streams(n, world) for s in lst)
            assert owned == list(range(n))
    assert parallel.env_rank()[1] >= 1
