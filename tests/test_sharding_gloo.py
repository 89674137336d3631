"""N > 1 host logic on CPU: world_size-2 (and 3) `gloo` process groups check that point-range sharding plus the
single all-reduce(MIN) of [min xyz, -max xyz] reproduces the AABB of the whole cloud.
The per-shard bounds come from the CPU oracle here (no GPU in this suite); on the GPU box the same vector is
produced by the fused convert kernel (tests/test_gpu_multigpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from pasture_b200 import sharding

N_POINTS = 10007


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, empty_rank, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = N_POINTS
        r = sharding.shard_range(n, rank, world)
        if rank == empty_rank:
            r = range(r.start, r.start)  # an empty shard must not disturb the result
        raw = O.gen_las_fmt0_records(r.start, len(r))
        ol_raw, ol_def = O.OLayout.las_raw(0), O.OLayout.las_default(0)
        src = O.OBuffer(ol_raw, len(r), False)
        if len(r):
            src.aos[:] = raw
        dst = O.OConverter.las_default(ol_raw, ol_def, (0.001,) * 3, (500000.0, 5400000.0, 100.0)).convert(src, True)
        b = O.calculate_bounds(dst)
        v = sharding.pack_bounds(None if b is None else (b[0], b[1]))
        sharding.allreduce_bounds(v)
        out_q.put((rank, v.tolist(), (r.start, r.stop)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,empty_rank", [(2, -1), (3, 1)])
def test_sharded_bounds_allreduce_equals_global(world, empty_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, empty_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: bounds over the union of the non-empty shards
    ol_raw, ol_def = O.OLayout.las_raw(0), O.OLayout.las_default(0)
    keep = []
    for rank, _, (a, b) in results:
        if rank != empty_rank:
            keep.append((a, b))
    mins, maxs = [], []
    for a, b in keep:
        src = O.OBuffer(ol_raw, b - a, False)
        src.aos[:] = O.gen_las_fmt0_records(a, b - a)
        dst = O.OConverter.las_default(ol_raw, ol_def, (0.001,) * 3, (500000.0, 5400000.0, 100.0)).convert(src, True)
        mn, mx = O.calculate_bounds(dst)
        mins.append(mn)
        maxs.append(mx)
    emin, emax = np.min(mins, axis=0), np.max(maxs, axis=0)
    for rank, v, _ in results:
        mn, mx = sharding.unpack_bounds(torch.tensor(v, dtype=torch.float64))
        assert list(mn) == list(emin) and list(mx) == list(emax), (rank, mn, mx)


def test_shard_ranges_partition_the_cloud():
    for n in (0, 1, 7, 100, 100_000_001):
        for world in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, r, world) for r in range(world)]
            assert rs[0].start == 0 and rs[-1].stop == n
            for a, b in zip(rs, rs[1:]):
                assert a.stop == b.start
            assert max(len(r) for r in rs) - min(len(r) for r in rs) <= 1
    assert sharding.unpack_bounds(sharding.pack_bounds(None)) is None
