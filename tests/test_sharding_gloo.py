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


# ---- sharded voxel grid: partials -> key-range all-to-all -> merge -------------------------------------------------
VOX_N, VOX_LEAF = 1500, (0.9, 1.1, 0.7)


def _vox_cloud():
    rng = np.random.default_rng(5)
    return rng.random((VOX_N, 3)) * [12.0, 9.0, 3.0] - [2.0, 4.0, 1.0]


def _vox_worker(rank, world, port, empty_rank, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = _vox_cloud()
        r = sharding.shard_range(VOX_N, rank, world)
        mine = pts[r.start:r.start] if rank == empty_rank else pts[r.start:r.stop]
        v = sharding.pack_bounds(None if len(mine) == 0 else (mine.min(0).tolist(), mine.max(0).tolist()))
        gmin, gmax = sharding.unpack_bounds(sharding.allreduce_bounds(v))
        k, c, s, bits, cells = O.voxel_partials(mine, gmin, gmax, VOX_LEAF)
        bounds = sharding.key_range_boundaries(cells[0], bits[1], bits[2], world)
        tk = torch.from_numpy(k)
        rk, rc, rs = sharding.exchange_partials(tk, torch.from_numpy(c), torch.from_numpy(s), sharding.split_sizes(tk, bounds))
        mk, mc, ms = O.merge_partials(rk.numpy(), rc.numpy(), rs.numpy())
        # every key this rank finalises lies in its own key range
        lo = 0 if rank == 0 else bounds[rank - 1]
        hi = bounds[rank] if rank < world - 1 else 1 << 62
        assert all(lo <= int(x) < hi for x in mk)
        out_q.put((rank, mk.tolist(), mc.tolist(), (ms / mc[:, None]).tolist(), bits))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,empty_rank", [(2, -1), (3, 0)])
def test_sharded_voxelgrid_equals_single_shot(world, empty_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vox_worker, args=(r, world, port, empty_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: the reference filter (oracle restatement) over the points that exist
    pts = _vox_cloud()
    if empty_rank >= 0:
        r = sharding.shard_range(VOX_N, empty_rank, world)
        pts = np.concatenate([pts[:r.start], pts[r.stop:]])
    ol = O.OLayout.from_attributes([("Position3D", O.VEC3F64)])
    ob = O.OBuffer(ol, len(pts), True)
    ob.set_attribute("Position3D", pts)
    oout, okeys = O.voxelgrid_filter(ob, VOX_LEAF, ol)
    bits = results[0][4]
    keys = np.array([k for _, ks, _, _, _ in results for k in ks], dtype=np.int64)  # rank order = key order
    cent = np.array([c for _, _, _, cs, _ in results for c in cs]).reshape(-1, 3)
    counts = np.array([c for _, _, cs, _, _ in results for c in cs])
    packed = (okeys[:, 0].astype(np.int64) << (bits[1] + bits[2])) | (okeys[:, 1].astype(np.int64) << bits[2]) | okeys[:, 2].astype(np.int64)
    assert np.array_equal(keys, packed)              # same voxels, same order
    assert counts.sum() == len(pts)
    expect = oout.attribute_bytes(0).view(np.float64).reshape(-1, 3)
    assert np.allclose(cent, expect, rtol=1e-9, atol=0)  # sum of per-shard sums vs one in-order sum


# ---- minmax_attribute over shards: the exchange and the seed rule (host logic; the per-shard fold is the GPU's job) ------

def _minmax_worker(rank, world, port, case, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pasture_b200 as pb
        name, np_dtype, attr, values = _minmax_case(case)
        r = sharding.shard_range(len(values), rank, world)
        if case.endswith("empty_first") and rank == 0:
            r = range(0, 0)
        if case.endswith("empty_first") and rank == 1:
            r = range(0, sharding.shard_range(len(values), 1, world).stop)
        shard = values[r.start:r.stop]
        local = first = None
        if len(shard):
            first = shard[0]
            with np.errstate(invalid="ignore"):
                local = (np.nanmin(shard, axis=0), np.nanmax(shard, axis=0)) if np.issubdtype(np_dtype, np.floating) \
                    else (shard.min(axis=0), shard.max(axis=0))
                if np.issubdtype(np_dtype, np.floating):  # a component without any non-NaN value: the fold's identity
                    local = (np.where(np.isnan(local[0]), sharding.F64_MAX, local[0]), np.where(np.isnan(local[1]), -sharding.F64_MAX, local[1]))
        res = sharding.combine_minmax(local, first, attr)
        out_q.put((rank, None if res is None else (np.asarray(res[0]).tolist(), np.asarray(res[1]).tolist())))
    finally:
        dist.destroy_process_group()


def _minmax_case(case):
    import pasture_b200 as pb
    from pasture_b200 import PointAttributeDefinition, PointAttributeDataType as DT
    rng = np.random.default_rng(11)
    base = case.replace("_empty_first", "")
    if base == "u64":
        v = rng.integers(0, 2 ** 64, 1001, dtype=np.uint64)
        v[500], v[3] = np.uint64(2 ** 64 - 1), np.uint64(0)
        return base, np.uint64, PointAttributeDefinition("PointID", DT.U64), v
    if base == "i8":
        return base, np.int8, PointAttributeDefinition("ScanAngleRank", DT.I8), rng.integers(-128, 128, 1001).astype(np.int8)
    if base == "vec3i32":
        return base, np.int32, PointAttributeDefinition("LASLocalPosition", DT.Vec3i32), rng.integers(-2 ** 31, 2 ** 31, (1001, 3)).astype(np.int32)
    if base == "f64_nan_later":
        v = rng.normal(size=1001)
        v[[1, 400, 1000]] = np.nan  # NaNs after the seed are ignored
        return base, np.float64, PointAttributeDefinition("GpsTime", DT.F64), v
    if base == "vec3f64_nan_seed":
        v = rng.normal(size=(1001, 3))
        v[0, 1] = np.nan  # a NaN seed sticks -- for that component only
        return base, np.float64, PointAttributeDefinition("Position3D", DT.Vec3f64), v
    raise KeyError(case)


def _sequential_minmax(values):
    """math/minmax.rs:62-96: seed = first value, strict comparisons"""
    v = np.asarray(values).reshape(len(values), -1)
    mn, mx = v[0].copy(), v[0].copy()
    for row in v[1:]:
        for c in range(v.shape[1]):
            if row[c] < mn[c]:
                mn[c] = row[c]
            if row[c] > mx[c]:
                mx[c] = row[c]
    return mn, mx


@pytest.mark.parametrize("case", ["u64", "i8", "vec3i32", "f64_nan_later", "vec3f64_nan_seed", "vec3f64_nan_seed_empty_first"])
def test_sharded_minmax_exchange_equals_the_sequential_fold(case):
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_minmax_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, _, _, values = _minmax_case(case)
    if case.endswith("empty_first"):  # rank 0 empty, rank 1 holds the cloud's start: the data ranks 0+1 would have held
        values = values
    emn, emx = _sequential_minmax(values)
    for rank, res in results:
        mn, mx = (np.atleast_1d(np.asarray(x, dtype=values.dtype)) for x in res)
        assert np.array_equal(mn, emn, equal_nan=True) and np.array_equal(mx, emx, equal_nan=True), (case, rank, mn, emn, mx, emx)


# ---- balanced key-range boundaries and the row exchange of the sharded voxel grid (host logic) ---------------------------

def _boundary_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        # a very uneven cloud: rank 0 sees only low keys, rank 1 a dense cluster, rank 2 nothing at all
        if rank == 0:
            keys = np.sort(rng.integers(0, 1 << 20, 5000))
        elif rank == 1:
            keys = np.sort(np.concatenate([rng.integers(1 << 30, (1 << 30) + 1000, 40000), rng.integers(0, 1 << 34, 3000)]))
        else:
            keys = np.zeros(0, dtype=np.int64)
        keys = torch.from_numpy(np.unique(keys).astype(np.int64))
        bounds = sharding.balanced_key_boundaries(keys, world)
        send = sharding.split_sizes(keys, bounds)
        recv = sharding._exchange_sizes(send, "cpu", None)
        rows = torch.stack([keys, keys * 2], dim=1)
        got = sharding._all_to_all_rows(rows, send, recv, None)
        out_q.put((rank, bounds, send, got[:, 0].tolist(), bool((got[:, 1] == got[:, 0] * 2).all())))
    finally:
        dist.destroy_process_group()


def test_balanced_boundaries_and_row_exchange():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_boundary_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = results[0][1]
    assert all(r[1] == bounds for r in results) and bounds == sorted(bounds) and len(bounds) == world - 1  # the same splitters everywhere
    assert all(r[4] for r in results)  # rows travel intact
    received = [r[3] for r in results]
    sizes = [len(x) for x in received]
    total = sum(sizes)
    assert total == sum(sum(r[2]) for r in results)
    assert max(sizes) < 0.6 * total  # the cluster is split up: no rank gets (nearly) everything, as equal key RANGES would give
    for d, keys in enumerate(received):  # ownership: rank d holds exactly the keys between its splitters
        lo = bounds[d - 1] if d > 0 else -1
        hi = bounds[d] if d < world - 1 else 1 << 62
        assert all(lo <= k < hi for k in keys) or d == 0 and all(k < hi for k in keys)
