"""Pins the oracle's BufferLayoutConverter + LAS default converter against the reference's LAS fixtures.

Golden input: tests/golden/las_fixtures.json (raw point records of pasture-io/resources/test/*.las, extracted by
tests/golden/make_las_golden.py).  Expected output: pasture-io/src/las/test_util.rs:46-183 (tests/las_expected.py)
and the different-layout test pasture-io/src/las/raw_readers.rs:815-911.
"""
import numpy as np
import pytest

import oracle as O
from tests import las_expected as E


def load_raw(entry, fmt, extra):
    raw = O.OLayout.las_raw(fmt)
    if extra:
        raw.add_attribute("ExtraBytesU32", O.U32, packed=1)  # las_layout.rs:167-169 (packed(1) extra bytes)
    assert raw.size == entry["record_length"]
    buf = O.OBuffer(raw, entry["count"], columnar=False)
    buf.aos[:] = np.frombuffer(bytes.fromhex(entry["records_hex"]), dtype=np.uint8)
    return raw, buf


@pytest.mark.parametrize("fmt", range(11))
@pytest.mark.parametrize("kind", ["plain", "extra_bytes"])
@pytest.mark.parametrize("columnar", [False, True])
def test_read_default_layout(las_fixtures, fmt, kind, columnar):
    entry = las_fixtures[kind][str(fmt)]
    raw, src = load_raw(entry, fmt, kind == "extra_bytes")
    target = O.OLayout.las_default(fmt)
    cv = O.OConverter.las_default(raw, target, entry["scale"], entry["offset"])
    dst = cv.convert(src, columnar)
    for name, expect in E.expected_default_layout_values(fmt).items():
        got = dst.attribute(name)
        assert got.dtype == expect.dtype, name
        assert np.array_equal(got, expect), (name, got, expect)


@pytest.mark.parametrize("fmt", range(11))
@pytest.mark.parametrize("columnar", [False, True])
def test_read_different_layout(las_fixtures, fmt, columnar):
    """raw_readers.rs:815-911: Vec3f32 positions, U32 classification, Vec3u8 colours (wrapping), zero-filled missing attrs"""
    entry = las_fixtures["plain"][str(fmt)]
    raw, src = load_raw(entry, fmt, False)
    target = O.OLayout.from_attributes([("Position3D", O.VEC3F32), ("Classification", O.U32),
                                        ("ColorRGB", O.VEC3U8), ("PointSourceID", O.U16),
                                        ("WaveformParameters", O.VEC3F32)])
    cv = O.OConverter.las_default(raw, target, entry["scale"], entry["offset"])
    dst = cv.convert(src, columnar)
    fl = E.fmt_flags(fmt)
    assert np.array_equal(dst.attribute("Position3D"), E.POSITIONS.astype(np.float32))
    assert np.array_equal(dst.attribute("Classification"), E.CLASSIFICATIONS.astype(np.uint32))
    exp_col = E.COLORS.astype(np.uint8) if fl["color"] else np.zeros((10, 3), np.uint8)
    assert np.array_equal(dst.attribute("ColorRGB"), exp_col)
    assert np.array_equal(dst.attribute("PointSourceID"), E.POINT_SOURCE_IDS)
    exp_w = E.WAVEPACKET_PARAMETERS if fl["waveform"] else np.zeros((10, 3), np.float32)
    assert np.array_equal(dst.attribute("WaveformParameters"), exp_w)


def test_read_in_chunks_matches_whole(las_fixtures):
    """raw_readers.rs:333-349: per-chunk convert_into_range into target sub-ranges"""
    entry = las_fixtures["plain"]["3"]
    raw, src = load_raw(entry, 3, False)
    target = O.OLayout.las_default(3)
    cv = O.OConverter.las_default(raw, target, entry["scale"], entry["offset"])
    whole = cv.convert(src, True)
    dst = O.OBuffer(target, 10, True)
    for b in range(0, 10, 3):
        e = min(b + 3, 10)
        chunk = O.OBuffer(raw, e - b, False)
        chunk.aos[:] = src.aos[b * raw.size: e * raw.size]
        cv.convert_into_range(chunk, 0, e - b, dst, b, e)
    for i in range(target.n):
        assert np.array_equal(whole.attribute_bytes(i), dst.attribute_bytes(i))


def test_invalid_position_type_is_error():
    raw = O.OLayout.las_raw(0)
    target = O.OLayout.from_attributes([("Position3D", O.VEC3I32)])
    with pytest.raises(O.OracleError) as e:
        O.OConverter.las_default(raw, target, [1, 1, 1], [0, 0, 0])  # raw_readers.rs:56 bail!
    assert e.value.code == O.ERR_UNSUPPORTED


def test_write_position_roundtrip_property():
    """pasture-io/tests/las_io.rs:245-350 + tests/common/mod.rs:56-78: integer coordinates in +-1000 at scale
    0.001 survive world -> i32 -> world exactly; the transform truncates (write_helpers.rs:15-17)."""
    import ctypes as C
    L = O.lib()
    scale = (C.c_double * 3)(0.001, 0.001, 0.001)
    off = (C.c_double * 3)(0.0, 0.0, 0.0)
    out = (C.c_int32 * 3)()
    for p in range(-1000, 1000):
        w = (C.c_double * 3)(float(p), float(p), float(p))
        assert L.po_las_write_position(w, scale, off, out) == 0
        back = out[0] * 0.001 + 0.0
        assert back == float(p)
    # truncation toward zero, not rounding
    w = (C.c_double * 3)(0.0019, -0.0019, 0.0005)
    L.po_las_write_position(w, scale, off, out)
    assert list(out) == [1, -1, 0]
    # out of i32 range -> the reference panics
    w = (C.c_double * 3)(3e6, 0.0, 0.0)
    assert L.po_las_write_position(w, scale, off, out) == 1
    w = (C.c_double * 3)(float("nan"), 0.0, 0.0)
    assert L.po_las_write_position(w, scale, off, out) == 0 and out[0] == 0


def test_write_direction_through_converter_matches_write_helper():
    """C1 configuration: POSITION_3D -> LASLocalPosition with INV_SCALE_OFFSET before the cast equals
    write_position_as_las_position for in-range values."""
    import ctypes as C
    n = 4096
    src_layout = O.OLayout.las_default(0)
    dst_layout = O.OLayout.las_raw(0)
    offset = (500000.0, 5400000.0, 100.0)
    src = O.OBuffer(src_layout, n, False)
    src.aos[:] = O.gen_c1_points(0, n, 42, offset)
    cv = O.OConverter(src_layout, dst_layout, with_default=True)
    t = O.make_transform(O.T_INV_SCALE_OFFSET, s=(0.001, 0.001, 0.001), o=offset)
    cv.set_custom_mapping_with_transformation(("Position3D", O.VEC3F64), ("LASLocalPosition", O.VEC3I32),
                                              O.VEC3F64, t, True)
    dst = cv.convert(src, False)
    got = dst.attribute("LASLocalPosition")
    pos = src.attribute("Position3D")
    L = O.lib()
    scale = (C.c_double * 3)(0.001, 0.001, 0.001)
    off = (C.c_double * 3)(*offset)
    out = (C.c_int32 * 3)()
    n_round_differs = 0
    for i in range(n):
        w = (C.c_double * 3)(*pos[i])
        assert L.po_las_write_position(w, scale, off, out) == 0
        assert list(out) == list(got[i])
        n_round_differs += int(np.any(np.rint((pos[i] - np.array(offset)) / 0.001) != got[i]))
    assert n_round_differs > n // 4  # truncation is observable on this stream
    assert np.array_equal(dst.attribute("Intensity"), src.attribute("Intensity"))
