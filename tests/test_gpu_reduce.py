"""GPU parity of the reductions (K5/K6/K10) against the oracle: calculate_bounds (bounds.rs:11-85),
minmax_attribute (minmax.rs:13-51), Morton codes on expand_bits_by_3 (bitmanip.rs:2-10)."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import attributes as A, PointAttributeDefinition, HashMapBuffer, VectorBuffer
from pasture_b200 import PointAttributeDataType as DT
from pasture_b200.algorithms import calculate_bounds, minmax_attribute, morton_codes
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("columnar", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 1000, 100003])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_bounds_default_positions(columnar, n, device):
    ol, pl = util.layouts([("Intensity", O.U16), ("Position3D", O.VEC3F64)], packed=1 if not columnar else 0)
    ob, pbuf = util.random_bytes_buffers(ol, pl, n, columnar, seed=n, device=device)
    omn, omx = O.calculate_bounds(ob)
    aabb = calculate_bounds(pbuf)
    assert list(aabb.min()) == list(omn) and list(aabb.max()) == list(omx)


def test_bounds_kats_and_none_cases():
    ol, pl = util.layouts([("Position3D", O.VEC3F64)])
    b = HashMapBuffer(pl, 3, "cuda")
    b.set_attribute("Position3D", [[0, 0, 0], [1, 1, 1], [-1, -1, -1]])  # math/bounds.rs:304-315
    aabb = calculate_bounds(b)
    assert aabb.min() == (-1.0, -1.0, -1.0) and aabb.max() == (1.0, 1.0, 1.0)
    assert calculate_bounds(HashMapBuffer(pl, 0, "cuda")) is None  # bounds.rs:12-14
    _, pl2 = util.layouts([("Intensity", O.U16)])
    assert calculate_bounds(HashMapBuffer(pl2, 4, "cuda")) is None  # :15-21
    b.set_attribute("Position3D", [[1, np.nan, 3], [0, 5, np.nan], [2, 4, 1]])  # NaN ignored :34-51
    aabb = calculate_bounds(b)
    assert aabb.min() == (0.0, 4.0, 1.0) and aabb.max() == (2.0, 5.0, 3.0)
    b.set_attribute("Position3D", np.full((3, 3), np.nan))
    with pytest.raises(pb.PastureB200Error):  # AABB::from_min_max panics (math/bounds.rs:21-26)
        calculate_bounds(b)


@pytest.mark.parametrize("dtype", [O.VEC3U8, O.VEC3U16, O.VEC3F32, O.VEC3I32])
@pytest.mark.parametrize("columnar", [False, True])
def test_bounds_custom_position_types(dtype, columnar):  # bounds.rs:56-85 converting view
    ol, pl = util.layouts([("Classification", O.U8), ("Position3D", dtype)], packed=1)
    ob, pbuf = util.random_bytes_buffers(ol, pl, 5003, columnar, seed=int(dtype))
    omn, omx = O.calculate_bounds(ob)
    aabb = calculate_bounds(pbuf)
    assert list(aabb.min()) == list(omn) and list(aabb.max()) == list(omx)


@pytest.mark.parametrize("dtype", list(range(10)) + [O.VEC3U8, O.VEC3U16, O.VEC3F32, O.VEC3I32, O.VEC3F64])
@pytest.mark.parametrize("columnar", [False, True])
def test_minmax_attribute_all_types(dtype, columnar):
    ol, pl = util.layouts([("pad", O.U8), ("v", dtype), ("tail", O.U16)], packed=1)
    ob, pbuf = util.random_bytes_buffers(ol, pl, 20011, columnar, seed=50 + int(dtype))
    omn, omx = O.minmax_attribute(ob, "v", dtype)
    mn, mx = minmax_attribute(pbuf, PointAttributeDefinition("v", dtype))
    assert np.array_equal(np.atleast_1d(mn), omn) and np.array_equal(np.atleast_1d(mx), omx)


def test_minmax_contract():
    ol, pl = util.layouts([("Intensity", O.I16), ("GpsTime", O.F64)])
    b = HashMapBuffer(pl, 4, "cuda")
    b.set_attribute("Intensity", [5, -3, 7, 0])
    b.set_attribute("GpsTime", [np.nan, -2.5, 0.0, 9.0])
    assert minmax_attribute(b, PointAttributeDefinition("Intensity", DT.I16)) == (-3, 7)
    mn, mx = minmax_attribute(b, A.GPS_TIME)  # a NaN seed is never replaced (math/minmax.rs:62-96)
    assert np.isnan(mn) and np.isnan(mx)
    b.set_attribute("GpsTime", [1.5, np.nan, 0.0, 9.0])
    assert minmax_attribute(b, A.GPS_TIME) == (0.0, 9.0)
    with pytest.raises(pb.PastureB200Error) as e:
        minmax_attribute(b, A.CLASSIFICATION)  # minmax.rs:17-26
    assert e.value.code == -1
    with pytest.raises(pb.PastureB200Error):
        minmax_attribute(b, PointAttributeDefinition("Intensity", DT.I32))  # datatype mismatch panics
    assert minmax_attribute(HashMapBuffer(pl, 0, "cuda"), A.GPS_TIME) is None


def _expand_bits_by_3_np(v):
    """math/bitmanip.rs:2-10 on a whole array (the same mask sequence as the oracle's scalar po_expand_bits_by_3)"""
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    for shift, mask in ((32, 0x00FF00000000FFFF), (16, 0x00FF0000FF0000FF), (8, 0xF00F00F00F00F00F), (4, 0x30C30C30C30C30C3), (2, 0x1249249249249249)):
        v = (v | (v << np.uint64(shift))) & np.uint64(mask)
    return v


@pytest.mark.parametrize("n,columnar", [(50_000, True), (1_000_003, True), (200_001, False)])
def test_morton_codes(n, columnar):
    """EVERY code of the cloud against the restatement (quantise to 21 bits per axis inside the AABB, interleave with
    expand_bits_by_3); the vectorised interleave itself is checked against the oracle's scalar function first"""
    L = O.lib()
    probe = np.array([0, 1, 2, 0x155555, 0x1FFFFF, 0xABCDE, 12345, 0x100000], dtype=np.uint64)
    assert [int(x) for x in _expand_bits_by_3_np(probe)] == [L.po_expand_bits_by_3(int(x)) for x in probe]
    pts = O.gen_terrain_positions(0, n)
    ol, pl = util.layouts([("Position3D", O.VEC3F64)] if columnar else [("Intensity", O.U16), ("Position3D", O.VEC3F64)])
    b = (HashMapBuffer if columnar else VectorBuffer)(pl, n, "cuda")
    b.set_attribute("Position3D", pts)
    mn, mx = pts.min(0), pts.max(0)
    codes = morton_codes(b, mn, mx).cpu().numpy().view(np.uint64)
    scale = np.where(mx - mn > 0, 2097152.0 / (mx - mn), 0.0)
    q = np.minimum(np.floor((pts - mn) * scale), 2097151).astype(np.uint64)
    expect = (_expand_bits_by_3_np(q[:, 0]) << np.uint64(2)) | (_expand_bits_by_3_np(q[:, 1]) << np.uint64(1)) | _expand_bits_by_3_np(q[:, 2])
    assert np.array_equal(codes, expect)
    assert int(codes.max()) < (1 << 63)


@pytest.mark.parametrize("dtype", [O.U8, O.I16, O.U64, O.I64, O.F32, O.F64, O.VEC3I32, O.VEC3F64])
def test_minmax_partial_and_logical_shards(dtype):
    """pb200_minmax_attribute_partial (the fold for a shard that does not start the cloud: NaN never enters) and the
    sharded helper: three logical shards of one buffer, combined like the ranks of a process group would, must give the
    reference's sequential result -- including a NaN seed in shard 0 and NaNs opening the later shards"""
    from pasture_b200 import sharding
    ol, pl = util.layouts([("pad", O.U8), ("v", dtype)], packed=1)
    n = 9001
    ob, pbuf = util.random_bytes_buffers(ol, pl, n, True, seed=70 + int(dtype), finite_floats=True)
    attr = PointAttributeDefinition("v", dtype)
    vals = ob.attribute("v").copy()
    is_float = dtype in (O.F32, O.F64, O.VEC3F64)
    for nan_seed in ([False, True] if is_float else [False]):
        if is_float:
            vals = vals.copy()
            flat = vals.reshape(n, -1)
            flat[3000, 0] = np.nan            # opens shard 1
            flat[6000:6003, -1] = np.nan      # opens shard 2
            flat[0, 0] = np.nan if nan_seed else 1.0
            ob.set_attribute("v", vals)
            pbuf = util.to_pb(ob, pl)
        omn, omx = O.minmax_attribute(ob, "v", dtype)  # the sequential fold of the whole cloud
        partials = []
        for lo, hi in ((0, 3000), (3000, 6000), (6000, n)):
            shard = HashMapBuffer(pl, hi - lo, "cuda", columns=[c[lo * pl.at(i).size(): hi * pl.at(i).size()] for i, c in enumerate(pbuf.columns)])
            local = minmax_attribute(shard, attr, partial=True)
            v = np.asarray(ob.attribute("v")[lo:hi]).reshape(hi - lo, -1)
            with np.errstate(invalid="ignore"):
                emn = np.nanmin(v, axis=0) if is_float else v.min(axis=0)
                emx = np.nanmax(v, axis=0) if is_float else v.max(axis=0)
            assert np.array_equal(np.atleast_1d(local[0]), emn) and np.array_equal(np.atleast_1d(local[1]), emx)
            partials.append((local, shard.slice_first(attr)))
        # combine as combine_minmax does across ranks: min / max of the partials, then the seed rule from shard 0
        mn = np.min([np.atleast_1d(p[0][0]) for p in partials], axis=0)
        mx = np.max([np.atleast_1d(p[0][1]) for p in partials], axis=0)
        if is_float:
            seed = np.atleast_1d(np.asarray(partials[0][1], dtype=np.float64))
            mn = np.where(np.isnan(seed), np.nan, mn)
            mx = np.where(np.isnan(seed), np.nan, mx)
        assert np.array_equal(mn, omn, equal_nan=True) and np.array_equal(mx, omx, equal_nan=True)
        # no process group: the helper degenerates to the single-device call
        r = sharding.minmax_attribute_sharded(pbuf, attr)
        assert np.array_equal(np.atleast_1d(r[0]), omn, equal_nan=True) and np.array_equal(np.atleast_1d(r[1]), omx, equal_nan=True)
