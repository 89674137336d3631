"""CPU-only checks of the product's boundary: the shared library loads, exports every symbol that
include/pasture_b200.h declares, the host-side PointLayout logic equals the oracle, and compute entry points
fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import _lib, attributes as A, PointLayout, PointAttributeDefinition, PointAttributeMember, FieldAlignment
from pasture_b200 import PointAttributeDataType as DT
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pasture_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.SO_PATH)
    names = declared_symbols()
    assert len(names) > 50
    for n in names:
        assert hasattr(L, n), f"{n} declared in pasture_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert L.pb200_abi_version() == 1


def test_layout_doc_asserts():
    l = PointLayout.from_attributes([A.POSITION_3D, A.INTENSITY])  # point_layout.rs:664-668, 924-927
    assert len(l) == 2 and l.at(0).offset() == 0 and l.at(1).offset() == A.POSITION_3D.size()
    assert l.size_of_point_entry() == 32
    l1 = PointLayout.from_attributes_packed([A.INTENSITY, A.POSITION_3D], 1)  # :684-691
    assert l1.at(0).offset() == 0 and l1.at(1).offset() == 2
    assert PointLayout.from_attributes_packed([A.INTENSITY, A.POSITION_3D], 4).at(1).offset() == 4
    l = PointLayout.from_members_and_alignment([A.INTENSITY.at_offset_in_type(2), A.POSITION_3D.at_offset_in_type(8)], 8)  # :713-717
    assert l.at(0).offset() == 2 and l.at(1).offset() == 8 and l.size_of_point_entry() == 32
    l = PointLayout.default()  # :770-776
    l.add_attribute(A.INTENSITY, FieldAlignment.Default)
    l.add_attribute(A.POSITION_3D, FieldAlignment.Default)
    assert l.at(0) == A.INTENSITY.at_offset_in_type(0) and l.at(1) == A.POSITION_3D.at_offset_in_type(8)
    assert l.has_attribute_with_name("Position3D") and l.has_attribute(A.POSITION_3D)  # :829, :845
    l.add_attribute(PointAttributeDefinition("X", DT.U32))
    assert not l.has_attribute(A.INTENSITY.with_custom_datatype(DT.U32))  # :848
    assert l.get_attribute(A.POSITION_3D.with_custom_datatype(DT.U32)) is None  # :865-866
    r = PointLayout.from_members_and_alignment([A.INTENSITY.at_offset_in_type(24), A.POSITION_3D.at_offset_in_type(0)], 8)  # :946-949
    assert r.index_of(A.INTENSITY) == 0 and r.index_of(A.POSITION_3D) == 1 and r.index_of(A.CLASSIFICATION) is None


def test_layout_errors():
    l = PointLayout.from_attributes([A.POSITION_3D])
    with pytest.raises(pb.PastureB200Error) as e:
        l.add_attribute(A.POSITION_3D.with_custom_datatype(DT.Vec3f32))
    assert e.value.code == -6
    with pytest.raises(pb.PastureB200Error) as e:
        PointLayout.from_members_and_alignment([PointAttributeMember.custom("a", DT.U32, 0), PointAttributeMember.custom("b", DT.U32, 2)], 4)
    assert e.value.code == -7


def test_las_layouts_match_oracle_and_reference_sizes():
    for f in range(11):
        util.las_layouts(f, True)
        util.las_layouts(f, False)
    assert [PointLayout.las_raw(f).size_of_point_entry() for f in range(11)] == [20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67]
    assert [PointLayout.las_default(f).size_of_point_entry() for f in range(11)] == [35, 43, 41, 49, 72, 78, 46, 52, 54, 75, 83]


def test_random_layouts_match_oracle():
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.integers(1, 12))
        attrs = [(f"a{i}", int(rng.integers(0, 16))) for i in range(n)]
        packed = int(rng.choice([0, 0, 1, 2, 4, 8]))
        util.layouts(attrs, packed)


def test_expand_bits_by_3_matches_oracle():
    rng = np.random.default_rng(0)
    for v in list(rng.integers(0, 1 << 63, 200)) + [0, 1, 2, 0x1FFFFF, 0x155555]:
        assert pb.algorithms.expand_bits_by_3(int(v)) == O.lib().po_expand_bits_by_3(int(v))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(pb.PastureB200Error) as e:
        pb.get_context()
    assert e.value.code == -100
    h = C.c_void_p()
    assert _lib.lib().pb200_ctx_create(0, C.byref(h)) == -100
    assert b"no CPU fallback" in _lib.lib().pb200_last_error()


def test_new_entry_points_reject_null_handles():
    """argument validation of the communicator / sort / sharded-voxel entry points happens before any CUDA call"""
    L = _lib.lib()
    h = C.c_void_p()
    INVALID = -8  # PB200_ERR_INVALID
    assert L.pb200_comm_create(None, 0, 1, C.byref(h)) == INVALID
    assert L.pb200_comm_handle(None, None) == INVALID
    assert L.pb200_comm_connect(None, None) == INVALID
    assert L.pb200_comm_check(None) == INVALID
    assert L.pb200_radix_sort_u64(None, None, None, 8, 0, 64) == INVALID
    assert L.pb200_voxelgrid_merge_partials(None, None, None, None, 0, 1, 1, 1, C.byref(h)) == INVALID
    assert L.pb200_voxel_partials_get(None, None) == INVALID
    assert L.pb200_converter_convert_into_range_with_global_bounds(None, None, 0, 0, None, 0, 0, None, None) < 0
    L.pb200_comm_destroy(None)            # destroying nothing is allowed
    L.pb200_voxel_partials_destroy(None)


def test_sharding_key_ranges_partition_every_key():
    """host logic of the key-range all-to-all: boundaries are ascending, every packed key has exactly one owner"""
    from pasture_b200 import sharding
    rng = np.random.default_rng(4)
    for cells_x, by, bz, world in [(1, 1, 1, 3), (10, 3, 2, 4), (5001, 13, 8, 8), (7, 4, 4, 2)]:
        b = sharding.key_range_boundaries(cells_x, by, bz, world)
        assert len(b) == world - 1 and b == sorted(b)
        keys = np.sort(rng.integers(0, max(1, cells_x) << (by + bz), 5000)).astype(np.int64)
        sizes = sharding.split_sizes(torch.from_numpy(keys), b)
        assert len(sizes) == world and sum(sizes) == len(keys)
        owner = np.searchsorted(np.array(b, dtype=np.int64), keys, side="right") if b else np.zeros(len(keys), int)
        assert [int((owner == d).sum()) for d in range(world)] == sizes


def test_product_does_not_import_oracle():
    """the product package must never route through oracle/"""
    pkg = os.path.join(ROOT, "pasture_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "pasture_oracle" not in text, os.path.join(dirpath, f)


def test_projection_pipeline_builder_host_side():
    """pb200_proj_pipeline_for_crs / pb200_proj_op_* are host-only: known CRS pairs, UTM zone parameters, refusals"""
    L = _lib.lib()
    ops = (_lib.ProjOp * 8)()
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", b"EPSG:3309", ops, 8) == 5
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", b"EPSG:3857", ops, 8) == 1 and ops[0].kind == 6
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:3857", b"EPSG:4326", ops, 8) == 1 and ops[0].kind == 11
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", b"EPSG:32633", ops, 8) == 2
    assert ops[0].kind == 8 and ops[1].kind == 7
    a, invf, lat0, lon0, k0, fe, fn = list(ops[1].p)[:7]
    assert (a, invf, lat0, k0, fe, fn) == (6378137.0, 298.257223563, 0.0, 0.9996, 500000.0, 0.0) and abs(np.degrees(lon0) - 15.0) < 1e-12
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", b"EPSG:32733", ops, 8) == 2 and ops[1].p[6] == 10000000.0
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", b"EPSG:25832", ops, 8) == 2 and ops[1].p[1] == 298.257222101
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:25832", b"EPSG:4326", ops, 8) == 2 and ops[0].kind == 10 and ops[1].kind == 9
    assert L.pb200_proj_pipeline_for_crs(b"EPSG:32632", b"EPSG:32633", ops, 8) == 2 and ops[0].kind == 10 and ops[1].kind == 7
    for bad in (b"EPSG:27700", b"EPSG:32661", b"EPSG:25839", b"+proj=utm +zone=32"):
        assert L.pb200_proj_pipeline_for_crs(b"EPSG:4326", bad, ops, 8) == -10
    op = _lib.ProjOp()
    assert L.pb200_proj_op_tmerc(6377563.396, 299.32496, 49.0, -2.0, 0.9996012717, 400000.0, -100000.0, 1, C.byref(op)) == 0
    assert op.kind == 10 and abs(op.p[2] - np.radians(49.0)) < 1e-15
    assert L.pb200_proj_op_tmerc(0.0, 299.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0, C.byref(op)) == -8
    assert L.pb200_proj_op_helmert(1.0, 2.0, 3.0, 0.0, 0.0, 0.0, 0.0, 0, C.byref(op)) == 0
    assert op.kind == 1 and list(op.p) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 2, 3]


def _schedule(cv_handle, src_layout, src_kind, dst_layout, dst_kind, n=1 << 20, sb=0, fresh=0):
    """pb200_converter_describe_schedule on fake device addresses (planning only: nothing is dereferenced)"""
    def desc(layout, kind, base):
        d = _lib.BufferDesc()
        d.layout, d.kind, d.memspace, d.len = layout._h, kind, 1, n
        cols = (C.c_void_p * len(layout))(*[base + (i + 1) * (1 << 28) for i in range(len(layout))])
        d.aos = C.c_void_p(base)
        d.columns = C.cast(cols, C.POINTER(C.c_void_p))
        d._keep = cols
        return d
    s, t = desc(src_layout, src_kind, 1 << 32), desc(dst_layout, dst_kind, 1 << 36)
    out = C.create_string_buffer(1 << 16)
    rc = _lib.lib().pb200_converter_describe_schedule(cv_handle, C.byref(s), sb, n, C.byref(t), 0, fresh, out, len(out))
    assert rc > 0, rc
    lines = out.value.decode().splitlines()
    head = dict(kv.split("=") for kv in lines[0].split()[1:])
    items = [dict(kv.split("=") for kv in l.split()[1:]) for l in lines[1:]]
    return lines[0].split()[0], head, items


def _check_schedule(head, items):
    """every point of the tile is covered exactly once per op, by at most threads/32 warps; grouped copies are cut at
    whole blocks of G x 32 points"""
    T, warps = int(head["tile_points"]), int(head["threads"]) // 32
    assert len(items) == int(head["items"]) <= int(head["ops"]) + warps
    assert max(int(i["warp"]) for i in items) < warps
    by_op = {}
    for i in items:  # an op is identified by what it reads and writes at point 0 of the tile
        p0, p1 = int(i["p0"]), int(i["p1"])
        assert 0 <= p0 < p1 <= T and p0 % 32 == 0 and (p1 % 32 == 0 or p1 == T)
        g = int(i["group"])
        if g:
            assert p0 % (32 * g) == 0 and (p1 % (32 * g) == 0 or p1 == T)
        by_op.setdefault((i["kind"], i["src_type"], i["dst_type"], i["xf"], i["bytes"]), []).append((p0, p1))
    n_rows = 0
    for key, cuts in by_op.items():
        total = sum(b - a for a, b in cuts)
        assert total % T == 0, (key, cuts)  # (several ops can share a key: x / y / z of a Vec3)
        n_rows += total // T
    assert n_rows == int(head["ops"])
    order = [int(i["warp"]) for i in items]
    assert order == sorted(order)  # the ops are laid out in order over the warps


def test_tile_schedule_host_side():
    """the converter's schedule (convert.cu: assign_items) is host logic: a planning-only converter (ctx == NULL) describes it
    without a device.  C2 (interleaved raw LAS fmt0 -> columnar default layout), the write direction with a fresh target, and
    columnar -> packed 35 B records (grouped copies)"""
    L = _lib.lib()
    raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    h = C.c_void_p()
    assert L.pb200_las_default_converter(None, raw._h, tgt._h, (C.c_double * 3)(0.001, 0.001, 0.001),
                                         (C.c_double * 3)(5e5, 5.4e6, 100.0), C.byref(h)) == 0
    mode, head, items = _schedule(h, raw, 0, tgt, 1)
    assert mode == "tiles" and int(head["tile_points"]) == 2048 and int(head["threads"]) == 512 and head["load_first"] == "1"
    assert int(head["ops"]) == 12
    _check_schedule(head, items)
    # a range that starts at an odd point: the tile shape stays, the streams are skewed
    mode2, head2, items2 = _schedule(h, raw, 0, tgt, 1, sb=3)
    _check_schedule(head2, items2)
    # conversions with a planning-only converter fail loudly
    d = _lib.BufferDesc()
    d.layout, d.kind, d.memspace, d.len, d.aos = raw._h, 0, 1, 16, C.c_void_p(1 << 32)
    e = _lib.BufferDesc()
    cols = (C.c_void_p * len(tgt))(*[(1 << 36) + i * 4096 for i in range(len(tgt))])
    e.layout, e.kind, e.memspace, e.len, e.columns = tgt._h, 1, 1, 16, C.cast(cols, C.POINTER(C.c_void_p))
    assert L.pb200_converter_convert_into_range(h, C.byref(d), 0, 16, C.byref(e), 0, 16, None) == -100
    L.pb200_converter_destroy(h)

    h = C.c_void_p()
    assert L.pb200_converter_create(None, tgt._h, tgt._h, 0, C.byref(h)) == 0
    mode, head, items = _schedule(h, tgt, 1, tgt, 0)  # columnar -> interleaved 35 B: every copy is grouped (G = 4)
    assert mode == "tiles" and all(i["kind"] == "copy" and i["group"] == "4" for i in items) and head["load_first"] == "0"
    _check_schedule(head, items)
    mode, head, items = _schedule(h, tgt, 0, tgt, 1)  # interleaved 35 B -> columnar: plain copies
    assert all(i["group"] == "0" for i in items)
    _check_schedule(head, items)
    L.pb200_converter_destroy(h)

    h = C.c_void_p()
    assert L.pb200_converter_create(None, tgt._h, raw._h, 1, C.byref(h)) == 0
    mode, head, items = _schedule(h, tgt, 0, raw, 0, fresh=1)  # write direction, `convert` semantics: zero ops fill the holes
    assert any(i["kind"] == "zero" for i in items)
    _check_schedule(head, items)
    mode, head, items = _schedule(h, tgt, 0, raw, 0, fresh=0)
    assert not any(i["kind"] == "zero" for i in items)
    L.pb200_converter_destroy(h)


@pytest.mark.parametrize("fmt", range(11))
def test_tile_schedule_of_every_las_format(fmt):
    """the schedule invariants (every op covers the tile exactly once, at most threads/32 warps, item limits) for the LAS
    read path (raw records -> default layout, both target kinds) and the write direction of every point format 0..10"""
    L = _lib.lib()
    raw, tgt = pb.PointLayout.las_raw(fmt), pb.PointLayout.las_default(fmt)
    h = C.c_void_p()
    assert L.pb200_las_default_converter(None, raw._h, tgt._h, (C.c_double * 3)(0.01, 0.01, 0.01),
                                         (C.c_double * 3)(0.0, 0.0, 0.0), C.byref(h)) == 0
    for dst_kind in (1, 0):
        mode, head, items = _schedule(h, raw, 0, tgt, dst_kind)
        assert mode == "tiles"
        _check_schedule(head, items)
    L.pb200_converter_destroy(h)
    h = C.c_void_p()
    assert L.pb200_converter_create(None, tgt._h, raw._h, 1, C.byref(h)) == 0
    for src_kind in (1, 0):
        for fresh in (0, 1):
            mode, head, items = _schedule(h, tgt, src_kind, raw, 0, fresh=fresh)
            assert mode == "tiles"
            _check_schedule(head, items)
    L.pb200_converter_destroy(h)


@pytest.mark.parametrize("seed", range(40))
def test_tile_schedule_of_random_layouts(seed):
    """random source / target layouts (all scalar and Vec3 types, byte arrays, packed and aligned members, missing and extra
    attributes), all four buffer-kind pairs, `convert` and `convert_into` semantics: the schedule invariants hold or the plan
    falls back to the direct kernel"""
    import numpy as np
    import oracle as O
    from tests import util
    rng = np.random.default_rng(9000 + seed)
    n_src = int(rng.integers(1, 12))
    src_attrs = [(f"a{i}", int(rng.choice(list(range(16)) + [O.BYTEARRAY])), int(rng.integers(1, 40))) for i in range(n_src)]
    src_attrs = [(a, d, e if d == O.BYTEARRAY else 0) for a, d, e in src_attrs]
    _, pl = util.layouts(src_attrs, packed=int(rng.choice([0, 1, 2])))
    dst_attrs = []
    for (a, d, e) in src_attrs:
        if rng.random() < 0.2:
            continue
        if d <= O.F64 and rng.random() < 0.6:
            d = int(rng.integers(0, 10))
        elif O.VEC3U8 <= d <= O.VEC3F64 and rng.random() < 0.6:
            d = int(rng.integers(O.VEC3U8, O.VEC3F64 + 1))
        dst_attrs.append((a, d, e))
    if rng.random() < 0.5:
        dst_attrs.append(("missing_in_source", O.U32, 0))
    if not dst_attrs:
        dst_attrs = [src_attrs[0]]
    _, plt = util.layouts(dst_attrs, packed=int(rng.choice([0, 1, 4])))
    L = _lib.lib()
    h = C.c_void_p()
    assert L.pb200_converter_create(None, pl._h, plt._h, 1, C.byref(h)) == 0
    for src_kind in (0, 1):
        for dst_kind in (0, 1):
            for fresh in (0, 1):
                mode, head, items = _schedule(h, pl, src_kind, plt, dst_kind, n=1 << 16, sb=int(rng.integers(0, 5)), fresh=fresh)
                if mode == "tiles":
                    _check_schedule(head, items)
    L.pb200_converter_destroy(h)
