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


def test_morton_codes():
    n = 50000
    pts = O.gen_terrain_positions(0, n)
    ol, pl = util.layouts([("Position3D", O.VEC3F64)])
    b = HashMapBuffer(pl, n, "cuda")
    b.set_attribute("Position3D", pts)
    mn, mx = pts.min(0), pts.max(0)
    codes = morton_codes(b, mn, mx).cpu().numpy().view(np.uint64)
    scale = np.where(mx - mn > 0, 2097152.0 / (mx - mn), 0.0)
    q = np.minimum(np.floor((pts - mn) * scale), 2097151).astype(np.uint64)
    L = O.lib()
    for i in list(range(50)) + [n - 1, int(np.argmax(pts[:, 0])), int(np.argmin(pts[:, 2]))]:
        e = (L.po_expand_bits_by_3(int(q[i, 0])) << 2) | (L.po_expand_bits_by_3(int(q[i, 1])) << 1) | L.po_expand_bits_by_3(int(q[i, 2]))
        assert int(codes[i]) == e
    assert int(codes.max()) < (1 << 63)
