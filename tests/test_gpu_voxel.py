"""GPU parity of the voxel-grid filter (K7-K9) against the oracle's faithful restatement of
pasture-algorithms/src/voxel_grid.rs (voxel ids, voxel count and order bit-exact; centroids bit-exact because
the stable sort keeps the reference's summation order; integer reductions exact, mode ties -> smallest value)."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import HashMapBuffer, VectorBuffer
from pasture_b200.algorithms import voxelgrid_filter
from tests import util
from tests.test_oracle_algorithms import COMPLETE, setup_point_cloud

pytestmark = pytest.mark.gpu


def to_product(ob, attrs, packed, device="cuda"):
    _, pl = util.layouts(attrs, packed)
    return util.to_pb(ob, pl, device), pl


@pytest.mark.parametrize("out_type", [HashMapBuffer, VectorBuffer])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_reference_voxel_test(out_type, device):  # voxel_grid.rs:906-938
    ol, ob = setup_point_cloud()
    pbuf, pl = to_product(ob, COMPLETE, 1, device)
    out, keys = voxelgrid_filter(pbuf, 1.0, 1.0, 1.0, out_buffer_type=out_type, return_keys=True)
    assert out.len() == 1000
    p = out.view_attribute("Position3D")[1]
    assert 0.59 < p[0] < 0.61 and 0.59 < p[1] < 0.61 and 1.59 < p[2] < 1.61
    assert out.view_attribute("Intensity")[1] == 4
    assert out.view_attribute("ReturnNumber")[1] == 42
    assert out.view_attribute("ClassificationFlags")[1] == 133
    oout, okeys = O.voxelgrid_filter(ob, (1.0, 1.0, 1.0), ol, columnar=True)
    assert np.array_equal(keys, okeys)
    util.assert_buffers_match(oout, out)


@pytest.mark.parametrize("seed,n,leaf", [(1, 5000, (0.7, 0.9, 1.1)), (2, 20000, (0.25, 0.25, 0.25)), (3, 777, (5.0, 0.01, 100.0)),
                                          (4, 1, (1.0, 1.0, 1.0)), (5, 3000, (100.0, 100.0, 100.0))])
@pytest.mark.parametrize("src_col", [False, True])
def test_random_clouds_match_faithful_oracle(seed, n, leaf, src_col):
    rng = np.random.default_rng(seed)
    attrs = [("Position3D", O.VEC3F64), ("Intensity", O.U16), ("Classification", O.U8), ("GpsTime", O.F64),
             ("ColorRGB", O.VEC3U16), ("ScanAngleRank", O.I8), ("ScanAngle", O.I16), ("PointSourceID", O.U16),
             ("PointID", O.U64), ("Normal", O.VEC3F32), ("EdgeOfFlightLine", O.U8), ("ClassificationFlags", O.U8)]
    ol = O.OLayout.from_attributes(attrs)
    ob = O.OBuffer(ol, n, src_col)
    ob.set_attribute("Position3D", rng.random((n, 3)) * [20, 10, 3] - [5, 5, 1])
    ob.set_attribute("Intensity", rng.integers(0, 65536, n))
    ob.set_attribute("Classification", rng.integers(0, 4, n))
    ob.set_attribute("GpsTime", rng.random(n) * 100 - 20)
    ob.set_attribute("ColorRGB", rng.integers(0, 65536, (n, 3)))
    ob.set_attribute("ScanAngleRank", rng.integers(-3, 3, n))
    ob.set_attribute("ScanAngle", rng.integers(-300, 300, n))
    ob.set_attribute("PointSourceID", rng.integers(0, 3, n))
    ob.set_attribute("PointID", rng.integers(0, 2**62, n))
    ob.set_attribute("Normal", rng.random((n, 3)).astype(np.float32) - 0.5)
    ob.set_attribute("EdgeOfFlightLine", rng.integers(0, 2, n))
    ob.set_attribute("ClassificationFlags", rng.integers(0, 256, n))
    pbuf, pl = to_product(ob, attrs, 0)
    oout, okeys = O.voxelgrid_filter(ob, leaf, ol, columnar=True, use_sort=False)
    out, keys = voxelgrid_filter(pbuf, *leaf, return_keys=True)
    assert out.len() == oout.len
    assert np.array_equal(keys, okeys)
    util.assert_buffers_match(oout, out)


def test_subset_target_layout_and_errors():
    ol, ob = setup_point_cloud(3)
    pbuf, pl = to_product(ob, COMPLETE, 1)
    _, sub = util.layouts([("Intensity", O.U16), ("Position3D", O.VEC3F64)])
    osub = O.OLayout.from_attributes([("Intensity", O.U16), ("Position3D", O.VEC3F64)])
    out = voxelgrid_filter(pbuf, 2.0, 2.0, 2.0, filtered_layout=sub)
    oout, _ = O.voxelgrid_filter(ob, (2.0, 2.0, 2.0), osub)
    util.assert_buffers_match(oout, out)
    _, bad = util.layouts([("Position3D", O.VEC3F64), ("WaveformPacketSize", O.U32)])
    with pytest.raises(pb.PastureB200Error) as e:  # voxel_grid.rs:452-459
        voxelgrid_filter(pbuf, 1.0, 1.0, 1.0, filtered_layout=bad)
    assert e.value.code == -10
    _, bad = util.layouts([("Position3D", O.VEC3F64), ("Custom", O.F32)])
    with pytest.raises(pb.PastureB200Error):  # :682-687
        voxelgrid_filter(pbuf, 1.0, 1.0, 1.0, filtered_layout=bad)
    _, nopos = util.layouts([("Position3D", O.VEC3F32)])
    with pytest.raises(pb.PastureB200Error) as e:  # :116-121
        voxelgrid_filter(HashMapBuffer(nopos, 4, "cuda"), 1.0, 1.0, 1.0)
    assert e.value.code == -1
    with pytest.raises(pb.PastureB200Error):  # empty buffer: calculate_bounds(..).unwrap()
        voxelgrid_filter(HashMapBuffer(pl, 0, "cuda"), 1.0, 1.0, 1.0)


def test_c3_shape_against_sort_oracle():
    """C3 stream (terrain positions, 0.1 m leaf) at a size the sort-based oracle restatement finishes in seconds"""
    n = 300000
    pts = O.gen_terrain_positions(0, n)
    ol = O.OLayout.from_attributes([("Position3D", O.VEC3F64)])
    ob = O.OBuffer(ol, n, True)
    ob.set_attribute("Position3D", pts)
    src = pb.algorithms.synth_terrain_positions(n)
    assert np.array_equal(src.view_attribute("Position3D"), pts)  # device generator == oracle generator, bit for bit
    oout, okeys = O.voxelgrid_filter(ob, (0.1, 0.1, 0.1), ol, columnar=True, use_sort=True)
    out, keys = voxelgrid_filter(src, 0.1, 0.1, 0.1, return_keys=True)
    assert out.len() == oout.len and np.array_equal(keys, okeys)
    util.assert_buffers_match(oout, out)


def test_fused_centroids_equal_the_per_voxel_kernel_at_scale():
    """the centroid computed inside the boundary pass (thread-per-point gather, in-order sums from shared memory, voxels
    running over tile ends) at sizes with tens of thousands of tiles: (1) position-only and position + other attributes
    (the path that also hands out segment starts and point indices) give the same bits, (2) both equal the oracle's
    sort-based restatement bit for bit on a 2 M-point prefix of the C3 stream, other reductions included"""
    n = 20_000_000
    src = pb.algorithms.synth_terrain_positions(n)
    out, keys = voxelgrid_filter(src, 0.1, 0.1, 0.1, return_keys=True)
    v = out.len()
    m = 2_000_000
    sub = HashMapBuffer(src.point_layout(), m, "cuda")
    sub.columns[0][: 24 * m].copy_(src.columns[0][: 24 * m])
    o2, k2 = voxelgrid_filter(sub, 0.1, 0.1, 0.1, return_keys=True)
    _, pl = util.layouts([("Position3D", O.VEC3F64), ("Intensity", O.U16), ("Classification", O.U8)])
    buf = HashMapBuffer(pl, m, "cuda")
    buf.columns[0][: 24 * m].copy_(src.columns[0][: 24 * m])
    g = torch.Generator(device="cuda").manual_seed(5)
    buf.columns[1][: 2 * m].copy_(torch.randint(0, 256, (2 * m,), dtype=torch.uint8, device="cuda", generator=g))
    buf.columns[2][:m].copy_(torch.randint(0, 5, (m,), dtype=torch.uint8, device="cuda", generator=g))
    o3, k3 = voxelgrid_filter(buf, 0.1, 0.1, 0.1, return_keys=True)  # index hand-out path (EMIT_INDEX) + other reductions
    assert np.array_equal(k2, k3) and torch.equal(o2.columns[0][: o2.len() * 24], o3.columns[0][: o3.len() * 24])
    # oracle on the same prefix: keys and centroids bit for bit, intensity mean / classification mode exact
    ol = O.OLayout.from_attributes([("Position3D", O.VEC3F64), ("Intensity", O.U16), ("Classification", O.U8)])
    ob = O.OBuffer(ol, m, True)
    ob.set_attribute("Position3D", src.columns[0][: 24 * m].cpu().numpy().view(np.float64).reshape(m, 3))
    ob.set_attribute("Intensity", buf.columns[1][: 2 * m].cpu().numpy().view(np.uint16))
    ob.set_attribute("Classification", buf.columns[2][:m].cpu().numpy())
    oout, okeys = O.voxelgrid_filter(ob, (0.1, 0.1, 0.1), ol, columnar=True, use_sort=True)
    assert np.array_equal(k3, okeys)
    util.assert_buffers_match(oout, o3)
    assert v > 0


def test_full_size_properties():
    """100 M-point C3 cloud: size-independent properties (keys strictly increasing, every point accounted for,
    centroids inside their voxel's marker neighbourhood, idempotent count)"""
    n = 100_000_000 if torch.cuda.get_device_properties(0).total_memory > 100e9 else 5_000_000
    src = pb.algorithms.synth_terrain_positions(n)
    out, keys = voxelgrid_filter(src, 0.1, 0.1, 0.1, return_keys=True)
    v = out.len()
    assert 0 < v <= n
    k = keys.astype(np.int64)
    packed = (k[:, 0] << 42) | (k[:, 1] << 21) | k[:, 2]
    assert np.all(np.diff(packed) > 0)  # lexicographic, unique
    c = out.view_attribute("Position3D")
    aabb = pb.algorithms.calculate_bounds(src)
    mn, mx = np.array(aabb.min()), np.array(aabb.max())
    assert np.all(c >= mn) and np.all(c <= mx)
    # nearest-marker rule: centroid index is within one cell of its voxel index
    approx = (c - mn) / 0.1
    assert np.all(np.abs(approx - keys.astype(np.float64)) < 1.6)
    # filtering the centroids again with the same grid origin keeps at most v voxels
    out2 = voxelgrid_filter(out, 0.1, 0.1, 0.1)
    assert out2.len() <= v


@pytest.mark.parametrize("leaf", [(50.0, 50.0, 50.0), (4.0, 4.0, 50.0)])
def test_crowded_voxels_use_the_sort_based_mode(leaf):
    """voxels with hundreds / thousands of points: the mode attributes take the sort-based path (composite-key radix
    sort + run-length vote) and must still equal the oracle (ties -> smallest value, signed values ordered)"""
    rng = np.random.default_rng(17)
    n = 6000
    attrs = [("Position3D", O.VEC3F64), ("Classification", O.U8), ("ScanAngleRank", O.I8), ("ScanAngle", O.I16),
             ("PointSourceID", O.U16), ("EdgeOfFlightLine", O.U8), ("Intensity", O.U16)]
    ol = O.OLayout.from_attributes(attrs)
    ob = O.OBuffer(ol, n, True)
    ob.set_attribute("Position3D", rng.random((n, 3)) * [20, 10, 3])
    ob.set_attribute("Classification", rng.integers(0, 5, n))
    ob.set_attribute("ScanAngleRank", rng.integers(-4, 4, n))
    ob.set_attribute("ScanAngle", rng.integers(-30000, 30000, n) // 10000 * 10000)
    ob.set_attribute("PointSourceID", rng.choice([0, 1, 40000, 65535], n))
    ob.set_attribute("EdgeOfFlightLine", rng.integers(0, 2, n))
    ob.set_attribute("Intensity", rng.integers(0, 65536, n))
    pbuf, _ = to_product(ob, attrs, 0)
    oout, okeys = O.voxelgrid_filter(ob, leaf, ol, columnar=True)
    out, keys = voxelgrid_filter(pbuf, *leaf, return_keys=True)
    assert np.array_equal(keys, okeys)
    util.assert_buffers_match(oout, out)


# ---- sharded voxel grid: S logical shards on one GPU + the exchange emulated by slicing ------------------
def _pos_cloud(pts, device="cuda"):
    ol, pl = util.layouts([("Position3D", O.VEC3F64)])
    ob = O.OBuffer(ol, len(pts), True)
    if len(pts):
        ob.set_attribute("Position3D", pts)
    return ob, ol, util.to_pb(ob, pl, device)


@pytest.mark.parametrize("shards,world", [(1, 1), (2, 2), (3, 2), (5, 4)])
def test_sharded_partials_and_merge_equal_single_shot(shards, world):
    from pasture_b200 import sharding
    from pasture_b200.algorithms import calculate_bounds, voxelgrid_merge_partials, voxelgrid_partials
    rng = np.random.default_rng(shards * 10 + world)
    n, leaf = 4000, (0.8, 1.3, 0.6)
    pts = rng.random((n, 3)) * [15.0, 8.0, 2.5] - [3.0, 1.0, 0.5]
    ob, ol, whole = _pos_cloud(pts)
    single, skeys = voxelgrid_filter(whole, *leaf, return_keys=True)
    b = calculate_bounds(whole)
    parts = []
    for r in range(shards):
        rr = sharding.shard_range(n, r, shards)
        _, _, shard = _pos_cloud(pts[rr.start:rr.stop])
        p = voxelgrid_partials(shard, *leaf, b)
        ok, oc, osum, obits, ocells = O.voxel_partials(pts[rr.start:rr.stop], b.min(), b.max(), leaf)
        assert list(p.bits) == list(obits) and list(p.cells) == list(ocells)
        assert np.array_equal(p.keys.cpu().numpy(), ok) and np.array_equal(p.counts.cpu().numpy(), oc)
        assert np.array_equal(p.sums.cpu().numpy(), osum)  # in-order sums: bit-exact
        parts.append(p)
    bounds = sharding.key_range_boundaries(parts[0].cells[0], parts[0].bits[1], parts[0].bits[2], world)
    sizes = [sharding.split_sizes(p.keys, bounds) for p in parts]
    out_keys, out_cent, out_counts = [], [], []
    for d in range(world):  # what destination rank d receives: every source's slice, in source order
        ks, cs, ss = [], [], []
        for p, sz in zip(parts, sizes):
            lo = sum(sz[:d])
            ks.append(p.keys[lo:lo + sz[d]]); cs.append(p.counts[lo:lo + sz[d]]); ss.append(p.sums[lo:lo + sz[d]])
        merged, cent = voxelgrid_merge_partials(torch.cat(ks), torch.cat(cs), torch.cat(ss), parts[0].bits, parts[0].cells)
        mk, mc, ms = O.merge_partials(torch.cat(ks).cpu().numpy(), torch.cat(cs).cpu().numpy(), torch.cat(ss).cpu().numpy())
        assert np.array_equal(merged.keys.cpu().numpy(), mk) and np.array_equal(merged.counts.cpu().numpy(), mc)
        assert np.array_equal(merged.sums.cpu().numpy(), ms)
        out_keys.append(merged.unpack_keys()); out_cent.append(cent); out_counts.append(merged.counts)
    keys = torch.cat(out_keys).cpu().numpy().astype(np.uint64)
    cent = torch.cat(out_cent).cpu().numpy()
    assert np.array_equal(keys, skeys)                       # same voxels in the reference's output order
    assert int(torch.cat(out_counts).sum()) == n
    ref = single.view_attribute("Position3D")
    if shards == 1:
        assert np.array_equal(cent, ref)                     # one shard: the same in-order sums
    else:
        assert np.allclose(cent, ref, rtol=1e-9, atol=0)     # sum of per-shard sums: 1e-9 relative


def test_sharded_voxelgrid_single_process_and_empty():
    from pasture_b200 import sharding
    pts = np.random.default_rng(9).random((2500, 3)) * 6.0
    _, _, whole = _pos_cloud(pts)
    single, skeys = voxelgrid_filter(whole, 0.5, 0.5, 0.5, return_keys=True)
    part, cent = sharding.voxelgrid_filter_sharded(whole, 0.5, 0.5, 0.5)  # no process group: one rank owns everything
    assert np.array_equal(part.unpack_keys().cpu().numpy().astype(np.uint64), skeys)
    assert np.array_equal(cent.cpu().numpy(), single.view_attribute("Position3D"))
    _, _, empty = _pos_cloud(np.zeros((0, 3)))
    assert sharding.voxelgrid_filter_sharded(empty, 0.5, 0.5, 0.5) is None
    b = pb.algorithms.calculate_bounds(whole)
    p = pb.algorithms.voxelgrid_partials(empty, 0.5, 0.5, 0.5, b)  # an empty shard of a non-empty cloud
    assert p.len() == 0 and p.cells[0] > 0


# ---- sharded voxel grid with every attribute reduction (pb200_voxelgrid_partials_layout / _merge_partials_layout) ---------

def _slice_buffer(buf, pl, lo, hi):
    return HashMapBuffer(pl, hi - lo, "cuda", columns=[c[lo * pl.at(i).size(): hi * pl.at(i).size()] for i, c in enumerate(buf.columns)])


def _cat_partials(parts):
    from pasture_b200.algorithms import VoxelAttrPartials, VoxelPartials
    pos = VoxelPartials(torch.cat([p.pos.keys for p in parts]), torch.cat([p.pos.counts for p in parts]), torch.cat([p.pos.sums for p in parts]),
                        parts[0].pos.bits, parts[0].pos.cells)
    cols = torch.cat([p.columns for p in parts])
    modes = [(torch.cat([p.modes[a][0] for p in parts]), torch.cat([p.modes[a][1] for p in parts])) for a in range(len(parts[0].modes))]
    return VoxelAttrPartials(pos, cols, parts[0].column_is_max, modes)


@pytest.mark.parametrize("seed,n,leaf,shards", [(1, 30000, (0.7, 0.9, 1.1), 3), (2, 50000, (0.25, 0.25, 0.25), 4), (3, 3002, (1.0, 1.0, 1.0), 2),
                                                 (4, 20000, (40.0, 40.0, 40.0), 5)])
def test_sharded_attribute_partials_merge_to_the_single_device_result(seed, n, leaf, shards):
    """S logical shards on one GPU: per-shard partials of ALL attributes (columns for the mean / max-pool rules, run lists
    for the "most common value" rules), concatenated in rank order and merged, must reproduce the single-device filter:
    voxel keys, counts and every integer attribute exactly (u16 sums are exact in f64, mode = merged histogram arg-max
    with ties to the smallest value, max-pool = max of maxima), f64 / f32 means within 1e-9 relative (the f64
    summation order differs).  leaf 40: a handful of crowded voxels; shards of unequal size, one of them empty."""
    from pasture_b200.algorithms import calculate_bounds, voxelgrid_merge_partials_layout, voxelgrid_partials_layout
    rng = np.random.default_rng(seed)
    attrs = [("Position3D", O.VEC3F64), ("Intensity", O.U16), ("Classification", O.U8), ("GpsTime", O.F64), ("ColorRGB", O.VEC3U16),
             ("ScanAngleRank", O.I8), ("ScanAngle", O.I16), ("PointSourceID", O.U16), ("PointID", O.U64), ("Normal", O.VEC3F32),
             ("EdgeOfFlightLine", O.U8), ("ClassificationFlags", O.U8), ("ReturnNumber", O.U8), ("NIR", O.U16)]
    ol = O.OLayout.from_attributes(attrs)
    ob = O.OBuffer(ol, n, True)
    ob.set_attribute("Position3D", rng.random((n, 3)) * [20, 10, 3] - [5, 5, 1])
    ob.set_attribute("Intensity", rng.integers(0, 65536, n))
    ob.set_attribute("Classification", rng.integers(0, 4, n))
    ob.set_attribute("GpsTime", rng.random(n) * 100 - 20)
    ob.set_attribute("ColorRGB", rng.integers(0, 65536, (n, 3)))
    ob.set_attribute("ScanAngleRank", rng.integers(-3, 3, n))
    ob.set_attribute("ScanAngle", rng.integers(-300, 300, n))
    ob.set_attribute("PointSourceID", rng.integers(0, 3, n))
    ob.set_attribute("PointID", rng.integers(0, 2**52, n))
    ob.set_attribute("Normal", rng.random((n, 3)).astype(np.float32) - 0.5)
    ob.set_attribute("EdgeOfFlightLine", rng.integers(0, 2, n))
    ob.set_attribute("ClassificationFlags", rng.integers(0, 256, n))
    ob.set_attribute("ReturnNumber", rng.integers(0, 8, n))
    ob.set_attribute("NIR", rng.integers(0, 65536, n))
    pbuf, pl = to_product(ob, attrs, 0)
    single, skeys = voxelgrid_filter(pbuf, *leaf, return_keys=True)
    b = calculate_bounds(pbuf)
    cuts = sorted(rng.integers(0, n, shards - 1).tolist())
    cuts[0] = cuts[1] if shards > 2 else cuts[0]  # an empty shard in the middle
    edges = [0] + cuts + [n]
    parts = [voxelgrid_partials_layout(_slice_buffer(pbuf, pl, lo, hi), *leaf, b, pl) for lo, hi in zip(edges[:-1], edges[1:])]
    assert sum(int(p.pos.counts.sum().item()) for p in parts) == n
    merged, mkeys = voxelgrid_merge_partials_layout(_cat_partials(parts), pl, return_keys=True)
    assert merged.len() == single.len() and np.array_equal(mkeys, skeys)
    for i, (name, dt) in enumerate(attrs):
        a, c = merged.view_attribute(name), single.view_attribute(name)
        if dt in (O.VEC3F64, O.VEC3F32):
            np.testing.assert_allclose(a, c, rtol=1e-6 if dt == O.VEC3F32 else 1e-9, atol=1e-12, err_msg=name)
        else:
            assert np.array_equal(a, c), name
    # a target layout without positions and with a subset of the attributes
    _, sub = util.layouts([("Classification", O.U8), ("Intensity", O.U16)])
    parts = [voxelgrid_partials_layout(_slice_buffer(pbuf, pl, lo, hi), *leaf, b, sub) for lo, hi in zip(edges[:-1], edges[1:])]
    m2 = voxelgrid_merge_partials_layout(_cat_partials(parts), sub)
    assert np.array_equal(m2.view_attribute("Classification"), single.view_attribute("Classification"))
    assert np.array_equal(m2.view_attribute("Intensity"), single.view_attribute("Intensity"))


def test_sharded_layout_helper_single_process():
    """no process group: sharding.voxelgrid_filter_sharded_layout == the single-device filter (all attributes)"""
    from pasture_b200 import sharding
    ol, ob = setup_point_cloud(5)
    pbuf, pl = to_product(ob, COMPLETE, 1)
    single, skeys = voxelgrid_filter(pbuf, 1.0, 1.0, 1.0, return_keys=True)
    t = {}
    out, keys = sharding.voxelgrid_filter_sharded_layout(pbuf, 1.0, 1.0, 1.0, return_keys=True, timings=t)
    assert np.array_equal(keys, skeys) and set(t) >= {"bounds+allreduce", "partials", "merge"}
    for i in range(len(pl)):
        name = pl.at(i).name()
        a, c = out.view_attribute(name), single.view_attribute(name)
        if a.dtype.kind == "f":
            np.testing.assert_allclose(a, c, rtol=1e-9, atol=1e-12, err_msg=name)
        else:
            assert np.array_equal(a, c), name
