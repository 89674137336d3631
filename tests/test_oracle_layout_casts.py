"""Pins the oracle's PointLayout rules and Rust-`as` casts against the reference's own asserts.

Layout: doc-test asserts of pasture-core/src/layout/point_layout.rs (:664-668, :684-691, :713-717,
:770-776, :911-912, :924-927, :946-949), LAS record sizes pasture-io/src/las/las_layout.rs:278 and
struct sizes pasture-io/src/las/las_types.rs:37..601.
Casts: Rust language `as` semantics behind attribute_conversion.rs:310-343 (edge vectors SURVEY App. A).
"""
import struct

import numpy as np
import pytest

import oracle as O

POSITION_3D = ("Position3D", O.VEC3F64)
INTENSITY = ("Intensity", O.U16)
COLOR_RGB = ("ColorRGB", O.VEC3U16)


def test_from_attributes_default_alignment():
    l = O.OLayout.from_attributes([POSITION_3D, INTENSITY])  # point_layout.rs:664-668, 924-927
    assert l.n == 2
    assert l.members()[0][2] == 0 and l.members()[1][2] == 24
    assert l.size == 32 and l.align == 8


def test_from_attributes_packed():
    l1 = O.OLayout.from_attributes([INTENSITY, POSITION_3D], packed=1)  # :684-687
    assert [m[2] for m in l1.members()] == [0, 2]
    l4 = O.OLayout.from_attributes([INTENSITY, POSITION_3D], packed=4)  # :689-691
    assert l4.members()[1][2] == 4


def test_add_attribute_default():
    l = O.OLayout.from_attributes([INTENSITY, POSITION_3D])  # :770-776
    assert [m[2] for m in l.members()] == [0, 8]


def test_from_members_and_alignment():
    l = O.OLayout.from_members_and_alignment([("Intensity", O.U16, 2), ("Position3D", O.VEC3F64, 8)], 8)  # :713-717
    assert [m[2] for m in l.members()] == [2, 8] and l.size == 32
    l2 = O.OLayout.from_members_and_alignment([("Intensity", O.U16, 24), ("Position3D", O.VEC3F64, 0)], 8)  # :947
    assert l2.index_by_name("Intensity") == 0 and l2.index_by_name("Position3D") == 1 and l2.size == 32
    with pytest.raises(O.OracleError):
        O.OLayout.from_members_and_alignment([("a", O.U32, 0), ("b", O.U32, 2)], 4)  # overlap :737-743
    with pytest.raises(O.OracleError):
        O.OLayout.from_members_and_alignment([("a", O.U32, 0), ("a", O.U32, 4)], 4)  # duplicate :725-730


def test_duplicate_attribute_panics():
    l = O.OLayout.from_attributes([POSITION_3D])
    with pytest.raises(O.OracleError) as e:
        l.add_attribute("Position3D", O.VEC3F32)
    assert e.value.code == O.ERR_DUPLICATE_ATTR


def test_derive_point_type_layout():
    # point_layout.rs:1045-1070 TestPoint1 repr(C, packed)
    l = O.OLayout.from_attributes([POSITION_3D, COLOR_RGB, INTENSITY], packed=1)
    assert [m[2] for m in l.members()] == [0, 24, 30] and l.size == 32 and l.align == 1


def test_las_record_sizes():
    expected_raw = [20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67]  # las_layout.rs:278
    expected_default = [35, 43, 41, 49, 72, 78, 46, 52, 54, 75, 83]  # las_types.rs const_assert_eq
    for f in range(11):
        assert O.OLayout.las_raw(f).size == expected_raw[f]
        assert O.OLayout.las_default(f).size == expected_default[f]
    raw0 = O.OLayout.las_raw(0)  # las_layout.rs:70-88
    assert raw0.members() == [("LASLocalPosition", O.VEC3I32, 0, 12), ("Intensity", O.U16, 12, 2),
                              ("LASBasicFlags", O.U8, 14, 1), ("Classification", O.U8, 15, 1),
                              ("ScanAngleRank", O.I8, 16, 1), ("UserData", O.U8, 17, 1),
                              ("PointSourceID", O.U16, 18, 2)]


def test_custom_point_types_test_utils():
    # pasture-core/src/test_utils.rs: CustomPointTypeSmall 25 B, CustomPointTypeBig 41 B (packed)
    small = O.OLayout.from_attributes([POSITION_3D, ("Classification", O.U8)], packed=1)
    big = O.OLayout.from_attributes([("GpsTime", O.F64), COLOR_RGB, POSITION_3D, ("Classification", O.U8),
                                     ("Intensity", O.I16)], packed=1)
    assert small.size == 25 and big.size == 41


def _cast(fr, to, fmt_from, value, fmt_to):
    out = O.convert_value(fr, to, struct.pack("<" + fmt_from, value))
    return struct.unpack("<" + fmt_to, out)[0]


def test_as_edge_vectors():
    nan, inf = float("nan"), float("inf")
    assert _cast(O.F64, O.U8, "d", -1.5, "B") == 0
    assert _cast(O.F64, O.U8, "d", -0.9, "B") == 0
    assert _cast(O.F64, O.U8, "d", 300.7, "B") == 255
    assert _cast(O.F64, O.I32, "d", 1e20, "i") == 2**31 - 1
    assert _cast(O.F64, O.I32, "d", -1e20, "i") == -2**31
    assert _cast(O.F64, O.I32, "d", nan, "i") == 0
    assert _cast(O.F64, O.U64, "d", nan, "Q") == 0
    assert _cast(O.F64, O.I64, "d", inf, "q") == 2**63 - 1
    assert _cast(O.F64, O.I64, "d", -inf, "q") == -2**63
    assert _cast(O.F64, O.U64, "d", 1.8446744073709552e19, "Q") == 2**64 - 1
    assert _cast(O.F64, O.U64, "d", 1.8446744073709550e19, "Q") == 18446744073709549568
    assert _cast(O.F32, O.I8, "f", 2.9, "b") == 2
    assert _cast(O.F32, O.I8, "f", -2.9, "b") == -2
    assert _cast(O.F32, O.I8, "f", 127.9, "b") == 127
    assert _cast(O.F32, O.I8, "f", -128.9, "b") == -128
    assert _cast(O.F32, O.U16, "f", 65535.9, "H") == 65535  # f32 rounds the literal to 65536.0 -> saturates
    assert _cast(O.I32, O.U8, "i", -1, "B") == 255
    assert _cast(O.U16, O.U8, "H", 0x1234, "B") == 0x34
    assert _cast(O.U8, O.I8, "B", 200, "b") == -56
    assert _cast(O.I8, O.U64, "b", -1, "Q") == 2**64 - 1
    assert _cast(O.I32, O.F32, "i", 16777217, "f") == 16777216.0
    assert _cast(O.U64, O.F64, "Q", 2**64 - 1, "d") == 18446744073709551616.0
    assert _cast(O.U64, O.F32, "Q", 2**64 - 1, "f") == 18446744073709551616.0
    assert _cast(O.I64, O.F32, "q", -(2**63), "f") == -9223372036854775808.0
    assert _cast(O.F64, O.F32, "d", 1e40, "f") == inf
    assert _cast(O.F64, O.F32, "d", 1e-50, "f") == 0.0
    assert _cast(O.F32, O.F64, "f", 0.1, "d") == struct.unpack("<f", struct.pack("<f", 0.1))[0]
    assert _cast(O.U32, O.I16, "I", 0xFFFF8001, "h") == -32767


def test_cast_table_shape():
    scalars = list(range(10))
    vec3 = [O.VEC3U8, O.VEC3U16, O.VEC3F32, O.VEC3I32, O.VEC3F64]
    n_s = sum(O.lib().po_has_conversion(a, b) for a in scalars for b in scalars)
    n_v = sum(O.lib().po_has_conversion(a, b) for a in vec3 for b in vec3)
    assert n_s == 90 and n_v == 20  # attribute_conversion.rs:194-260
    assert not O.lib().po_has_conversion(O.VEC4U8, O.VEC3U8)
    assert not O.lib().po_has_conversion(O.U8, O.VEC3U8)
    with pytest.raises(O.OracleError):
        O.convert_value(O.VEC4U8, O.U32, b"\0\0\0\0")


def test_scalar_casts_match_numpy_in_range():
    """in-range values: numpy astype == Rust `as` (wrap for ints, RNE for int->float)."""
    rng = np.random.default_rng(7)
    for fr, nf in O.NP_DTYPES.items():
        if np.issubdtype(nf, np.integer):
            info = np.iinfo(nf)
            vals = rng.integers(info.min, info.max, size=64, dtype=nf, endpoint=True)
        else:
            vals = (rng.random(64) * 200 - 100).astype(nf)
        for to, nt in O.NP_DTYPES.items():
            if fr == to:
                continue
            if np.issubdtype(nf, np.floating) and np.issubdtype(nt, np.integer):
                info = np.iinfo(nt)
                v = np.clip(np.trunc(vals.astype(np.float64)), max(info.min, -100), min(info.max, 100))
                expect = v.astype(nt)
            else:
                with np.errstate(over="ignore"):
                    expect = vals.astype(nt)
            for x, e in zip(vals, expect):
                got = np.frombuffer(O.convert_value(fr, to, np.array([x], dtype=nf).tobytes()), dtype=nt)[0]
                assert got == e, (fr, to, x, got, e)


def test_vec3_cast_componentwise():
    v = np.array([1.9, -2.9, 70000.5], dtype=np.float64)
    out = np.frombuffer(O.convert_value(O.VEC3F64, O.VEC3U16, v.tobytes()), dtype=np.uint16)
    assert list(out) == [1, 0, 65535]
    out = np.frombuffer(O.convert_value(O.VEC3F64, O.VEC3I32, v.tobytes()), dtype=np.int32)
    assert list(out) == [1, -2, 70000]
    c = np.array([0, 1 << 4, 2 << 8], dtype=np.uint16)
    out = np.frombuffer(O.convert_value(O.VEC3U16, O.VEC3U8, c.tobytes()), dtype=np.uint8)
    assert list(out) == [0, 16, 0]  # raw_readers.rs:866-872 `c as u8` wraps


def test_expand_bits_by_3():
    L = O.lib()
    assert L.po_expand_bits_by_3(1) == 1 and L.po_expand_bits_by_3(2) == 8
    assert L.po_expand_bits_by_3(0x1FFFFF) == 0x1249249249249249
    assert L.po_expand_bits_by_3(0x155555) == 0x1041041041041041
    assert L.po_expand_bits_by_3(0xFFFFFFFFFFFFFFFF) == 0x1249249249249249  # truncates to 21 bits first
    rng = np.random.default_rng(1)
    for v in rng.integers(0, 1 << 21, size=200):
        naive = 0
        for b in range(21):
            naive |= ((int(v) >> b) & 1) << (3 * b)
        assert L.po_expand_bits_by_3(int(v)) == naive
    assert L.po_reverse_bits(1) == 1 << 63 and L.po_reverse_bits(0xF0) == 0x0F << 56
