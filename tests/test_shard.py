"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo processes exercise the sharding + final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flope_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 255, 256, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(s for s in sizes if s or n == 0) <= max(sizes)      # contiguous, near-equal
            assert max(sizes) == (-(-n // world) if n else 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, micro, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def fake_pose_fn(lo, hi):           # stands in for engine.infer_frames on crops [lo, hi)
            idx = torch.arange(lo, hi, dtype=torch.float64)
            return torch.stack([idx, idx * idx, -idx], 1).reshape(-1, 3)
        out = shard.run_sharded(fake_pose_fn, n_total, micro)
        want = torch.arange(n_total, dtype=torch.float64)
        ok = out.shape == (n_total, 3) and torch.equal(out[:, 0], want) and torch.equal(out[:, 1], want * want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,micro", [(2, 37, 8), (2, 1, 4), (3, 10, 3), (2, 64, 64)])
def test_sharded_run_gathers_rows_in_original_order(world, n_total, micro):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, micro, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert all(res.values()) and len(res) == world
