"""world_size-2 gloo tests of the multi-GPU host logic: utterance sharding covers the batch
exactly once and the metric all-reduce reproduces the single-process batch means."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_sass_tf_b200.app import parallel


@pytest.mark.parametrize("n,world", [(256, 1), (256, 2), (8192, 8), (7, 4), (3, 8), (0, 2)])
def test_shard_range_partitions(n, world):
    seen = []
    for r in range(world):
        lo, hi = parallel.shard_range(n, r, world)
        assert 0 <= lo <= hi <= n
        seen += list(range(lo, hi))
    assert seen == list(range(n))
    sizes = [parallel.shard_range(n, r, world) for r in range(world)]
    assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        snr = rng.normal(10, 3, size=37)            # per-utterance metrics of the whole batch
        ae = rng.random(37)
        lo, hi = parallel.shard_range(37, rank, world)
        vec = parallel.metric_vector(snr[lo:hi].sum(), ae[lo:hi].sum(), 0.0, hi - lo)
        got = parallel.allreduce_metrics(vec)
        q.put((rank, got, (float(snr.mean()), float(ae.mean()))))
    finally:
        dist.destroy_process_group()


def test_metric_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, want in res:
        assert abs(got[0] - want[0]) < 1e-4 and abs(got[1] - want[1]) < 1e-5
        assert got[3] == 37.0


def test_allreduce_without_group_is_identity():
    v = parallel.metric_vector(6.0, 3.0, 0.0, 3.0)
    assert parallel.allreduce_metrics(v) == (2.0, 1.0, 0.0, 3.0)
    assert parallel.allreduce_metrics(parallel.metric_vector()) == (0.0, 0.0, 0.0, 0.0)
