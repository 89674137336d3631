"""Oracle BufferLayoutConverter vs the reference's own property tests (buffer_conversion.rs:684-930):
random CustomPointTypeBig points, every {Vector,HashMap} x {Vector,HashMap} pair, compared with the same
operation done through typed views (numpy here)."""
import itertools

import numpy as np
import pytest

import oracle as O

POSITION_3D = ("Position3D", O.VEC3F64)
CLASSIFICATION = ("Classification", O.U8)
RETURN_NUMBER = ("ReturnNumber", O.U8)
BIG = [("GpsTime", O.F64), ("ColorRGB", O.VEC3U16), POSITION_3D, CLASSIFICATION, ("Intensity", O.I16)]
SMALL = [POSITION_3D, CLASSIFICATION]
PAIRS = list(itertools.product([False, True], [False, True]))


def random_big(n, columnar, seed=0):
    rng = np.random.default_rng(seed)
    l = O.OLayout.from_attributes(BIG, packed=1)
    b = O.OBuffer(l, n, columnar)
    b.set_attribute("GpsTime", rng.random(n))
    b.set_attribute("ColorRGB", rng.integers(0, 65536, (n, 3)))
    b.set_attribute("Position3D", rng.random((n, 3)))
    b.set_attribute("Classification", rng.integers(0, 256, n))
    b.set_attribute("Intensity", rng.integers(-32768, 32768, n))
    return l, b


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_default(src_col, dst_col):  # :684-722
    l, src = random_big(16, src_col)
    target = O.OLayout.from_attributes(SMALL, packed=1)
    out = O.OConverter(l, target).convert(src, dst_col)
    assert np.array_equal(out.attribute("Position3D"), src.attribute("Position3D"))
    assert np.array_equal(out.attribute("Classification"), src.attribute("Classification"))


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_multiple_attributes_from_one(src_col, dst_col):  # :724-762
    l, src = random_big(16, src_col, 1)
    target = O.OLayout.from_attributes([CLASSIFICATION, RETURN_NUMBER])
    cv = O.OConverter(l, target, with_default=True)
    cv.set_custom_mapping(CLASSIFICATION, RETURN_NUMBER)
    out = cv.convert(src, dst_col)
    assert np.array_equal(out.attribute("Classification"), src.attribute("Classification"))
    assert np.array_equal(out.attribute("ReturnNumber"), src.attribute("Classification"))


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
@pytest.mark.parametrize("apply_to_source", [True, False])
def test_buffer_converter_transformed_attribute(src_col, dst_col, apply_to_source):  # :764-846
    l, src = random_big(16, src_col, 2)
    target = O.OLayout.from_attributes([POSITION_3D])
    cv = O.OConverter(l, target, with_default=True)
    cv.set_custom_mapping_with_transformation(POSITION_3D, POSITION_3D, O.VEC3F64,
                                              O.make_transform(O.T_ADD, o=(42.0, 42.0, 42.0)), apply_to_source)
    out = cv.convert(src, dst_col)
    assert np.array_equal(out.attribute("Position3D"), src.attribute("Position3D") + 42.0)


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_identity(src_col, dst_col):  # :848-873
    l, src = random_big(16, src_col, 3)
    out = O.OConverter(l, l, with_default=True).convert(src, dst_col)
    for i in range(l.n):
        assert np.array_equal(out.attribute_bytes(i), src.attribute_bytes(i))


def test_buffer_converter_mismatched_len():  # :912-930 should_panic
    l, src = random_big(16, False, 4)
    dst = O.OBuffer(l, 8, False)
    with pytest.raises(O.OracleError) as e:
        O.OConverter(l, l, with_default=True).convert_into(src, dst)
    assert e.value.code == O.ERR_RANGE


def test_missing_attribute_panics_without_default():  # :112-116
    l, _ = random_big(1, False)
    target = O.OLayout.from_attributes([RETURN_NUMBER])
    with pytest.raises(O.OracleError) as e:
        O.OConverter(l, target)
    assert e.value.code == O.ERR_ATTR_NOT_FOUND
    cv = O.OConverter(l, target, with_default=True)
    assert cv.c.n_mappings == 0


def test_transform_dtype_assert():  # :209-213
    l, _ = random_big(1, False)
    target = O.OLayout.from_attributes([("Position3D", O.VEC3F32)])
    cv = O.OConverter(l, target, with_default=True)
    t = O.make_transform(O.T_ADD, o=(1, 1, 1))
    with pytest.raises(O.OracleError) as e:  # transform after the cast must have the TARGET type
        cv.set_custom_mapping_with_transformation(POSITION_3D, ("Position3D", O.VEC3F32), O.VEC3F64, t, False)
    assert e.value.code == O.ERR_TRANSFORM_DTYPE
    cv.set_custom_mapping_with_transformation(POSITION_3D, ("Position3D", O.VEC3F32), O.VEC3F64, t, True)
    cv.set_custom_mapping_with_transformation(POSITION_3D, ("Position3D", O.VEC3F32), O.VEC3F32, t, False)


def test_impossible_cast_panics():  # :383-388
    a = O.OLayout.from_attributes([("X", O.VEC4U8)])
    b = O.OLayout.from_attributes([("X", O.U32)])
    with pytest.raises(O.OracleError) as e:
        O.OConverter(a, b)
    assert e.value.code == O.ERR_NO_CONVERSION


def test_layout_mismatch_panics():  # :302-303
    l, src = random_big(4, False)
    other = O.OLayout.from_attributes(BIG)  # default alignment: different offsets
    dst = O.OBuffer(other, 4, False)
    cv = O.OConverter(l, l, with_default=True)
    with pytest.raises(O.OracleError) as e:
        cv.convert_into(src, dst)
    assert e.value.code == O.ERR_LAYOUT_MISMATCH


def test_unmapped_target_bytes_untouched():
    l, src = random_big(8, False)
    target = O.OLayout.from_attributes([POSITION_3D, ("Unmapped", O.U32), CLASSIFICATION])
    dst = O.OBuffer(target, 8, False)
    dst.aos[:] = 0xAB
    O.OConverter(l, target, with_default=True).convert_into(src, dst)
    assert np.all(dst.attribute_bytes(1) == 0xAB)
    rec = dst.aos.reshape(8, target.size)
    assert np.all(rec[:, 29:] == 0xAB)  # tail padding
    assert np.array_equal(dst.attribute("Position3D"), src.attribute("Position3D"))


def test_mt_variant_equals_single_thread():
    n = 10007
    raw = O.OLayout.las_raw(0)
    src = O.OBuffer(raw, n, False)
    src.aos[:] = O.gen_las_fmt0_records(0, n)
    target = O.OLayout.las_default(0)
    cv = O.OConverter.las_default(raw, target, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
    a = cv.convert(src, True)
    b = O.OBuffer(target, n, True)
    cv.convert_into_range(src, 0, n, b, 0, n, threads=5)
    for i in range(target.n):
        assert np.array_equal(a.attribute_bytes(i), b.attribute_bytes(i))
    # numpy restatement of the C2 mapping (raw_readers.rs:42-48, :61-103)
    rec = src.aos.reshape(n, 20)
    xyz = np.ascontiguousarray(rec[:, :12]).view(np.int32).reshape(n, 3)
    pos = xyz.astype(np.float64) * 0.001 + np.array([500000.0, 5400000.0, 100.0])
    assert np.array_equal(a.attribute("Position3D"), pos)
    flags = rec[:, 14]
    assert np.array_equal(a.attribute("ReturnNumber"), flags & 7)
    assert np.array_equal(a.attribute("NumberOfReturns"), (flags >> 3) & 7)
    assert np.array_equal(a.attribute("ScanDirectionFlag"), (flags >> 6) & 1)
    assert np.array_equal(a.attribute("EdgeOfFlightLine"), (flags >> 7) & 1)
    assert np.array_equal(a.attribute("ScanAngleRank"), rec[:, 16].view(np.int8))


def test_oracle_filter_into_like_reference_test():  # point_buffer.rs:2296-2329 (every second point survives)
    ol = O.OLayout.from_attributes([("Position3D", O.VEC3F64), ("Classification", O.U8)])
    src = O.OBuffer(ol, 8, True)
    src.set_attribute("Position3D", np.arange(24, dtype=np.float64).reshape(8, 3))
    src.set_attribute("Classification", np.arange(8))
    for columnar in (True, False):
        dst = O.OBuffer(ol, 4, columnar)
        assert O.filter_into(src, lambda i: i % 2 == 0, dst) == 4
        assert np.array_equal(dst.attribute("Classification"), [0, 2, 4, 6])
        assert np.array_equal(dst.attribute("Position3D"), np.arange(24, dtype=np.float64).reshape(8, 3)[::2])
    with pytest.raises(O.OracleError):
        O.filter_into(src, lambda i: True, O.OBuffer(ol, 4, True))


@pytest.mark.parametrize("scale", [0.001, 0.01, -0.37, 3.0, 1e-300, 1e300, 5e-324, 123456.789])
def test_inverse_scale_offset_is_ieee_division_then_rust_cast(scale):
    """pins the oracle side of tests/test_gpu_convert.py::test_inverse_scale_offset_division_bit_exact: `(v - offset) / scale`
    (write_helpers.rs:15-17) in the oracle is the IEEE-754 subtraction and division numpy performs, followed by Rust's
    saturating `as i32` (NaN -> 0)"""
    rng = np.random.default_rng(int(abs(scale) * 1e6) % 1000 + 5)
    n = 200_000
    offset = 500000.0
    vals = rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64).copy()
    q = rng.integers(-2 ** 31 - 5, 2 ** 31 + 5, n // 2).astype(np.float64)
    with np.errstate(all="ignore"):
        vals[: n // 2] = q * scale + offset
    ol = O.OLayout.from_attributes([("v", O.F64, 0)])
    olt = O.OLayout.from_attributes([("v", O.I32, 0)])
    src = O.OBuffer(ol, n, True)
    src.columns[0][:] = vals.view(np.uint8)
    cv = O.OConverter(ol, olt, with_default=False)
    cv.set_custom_mapping_with_transformation(("v", O.F64), ("v", O.I32), O.F64,
                                               O.make_transform(O.T_INV_SCALE_OFFSET,
                                                                s=(scale,) * 3, o=(offset,) * 3), True)
    got = cv.convert(src, True).columns[0][: 4 * n].view(np.int32)
    with np.errstate(all="ignore"):
        t = (vals - offset) / scale
    want = np.where(np.isnan(t), 0.0, np.clip(np.trunc(t), -2.0 ** 31, 2.0 ** 31 - 1)).astype(np.int64).astype(np.int32)
    assert np.array_equal(got, want)
