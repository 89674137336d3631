"""The .pnts restatement (oracle.pnts_read_into / pnts_feature_table_body) against the reference's fixture
pasture-io/resources/test/points.pnts and its tests (pnts_reader.rs:420-482, pnts_writer.rs:448-597)."""
import json
import struct

import numpy as np
import pytest

import oracle as O
from tests.pnts_expected import check_fixture_arrays, fixture


def file_attrs(blob):
    ft_json = struct.unpack("<I", blob[12:16])[0]
    header = json.loads(blob[28:28 + ft_json])
    body = 28 + ft_json
    attrs = []
    for semantic, name, dtype in (("POSITION", "Position3D", O.VEC3F32), ("RGBA", "ColorRGBA", O.VEC4U8),
                                  ("RGB", "ColorRGB", O.VEC3U8), ("NORMAL", "Normal", O.VEC3F32)):
        if semantic in header:
            attrs.append((name, dtype, body + header[semantic]["byteOffset"]))
    return attrs, header


@pytest.mark.parametrize("columnar", [True, False])
def test_reference_fixture_default_layout(columnar):
    blob, exp = fixture()
    attrs, header = file_attrs(blob)
    n = header["POINTS_LENGTH"]
    assert n == exp["points_length"] == 8000
    dst = O.OBuffer(O.OLayout.from_attributes([(a[0], a[1]) for a in attrs], packed=1), n, columnar)
    O.pnts_read_into(blob, attrs, 0, n, dst)
    check_fixture_arrays(dst.attribute("Position3D"), dst.attribute("ColorRGB"), exp)


def test_reference_fixture_into_other_layout_and_window():
    blob, exp = fixture()
    attrs, _ = file_attrs(blob)
    layout = O.OLayout.from_attributes([("Intensity", O.U16), ("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16)])
    dst = O.OBuffer(layout, 100, False)
    dst.set_attribute("Intensity", np.arange(100))
    O.pnts_read_into(blob, attrs, 7900, 100, dst)
    assert dst.attribute("Position3D")[-4:].tolist() == exp["last_positions"]  # f32 -> f64 is exact
    assert dst.attribute("ColorRGB")[-4:].tolist() == exp["last_rgb"]
    assert np.array_equal(dst.attribute("Intensity"), np.arange(100))  # not in the file: left alone


def body_of(points, dtype):
    return np.asarray(points, dtype=dtype).tobytes()


@pytest.mark.parametrize("mode_absolute", [True, False])
def test_read_modes(mode_absolute):  # pnts_reader.rs:420-482
    body = body_of([[10, 10, 10], [20, 20, 20]], "<f4")
    dst = O.OBuffer(O.OLayout.from_attributes([("Position3D", O.VEC3F32)], packed=1), 2, False)
    O.pnts_read_into(body, [("Position3D", O.VEC3F32, 0)], 0, 2, dst, [10.0, 10.0, 10.0] if mode_absolute else None)
    want = [[20, 20, 20], [30, 30, 30]] if mode_absolute else [[10, 10, 10], [20, 20, 20]]
    assert dst.attribute("Position3D").tolist() == want


def test_rtc_is_added_in_f64_then_rounded_to_f32():  # pnts_reader.rs:265-273
    body = body_of([[0.1, 16777216.0, -3.3]], "<f4")
    dst = O.OBuffer(O.OLayout.from_attributes([("Position3D", O.VEC3F32)], packed=1), 1, True)
    c = [1e-9, 1.0, 3.3]
    O.pnts_read_into(body, [("Position3D", O.VEC3F32, 0)], 0, 1, dst, c)
    p = np.array([0.1, 16777216.0, -3.3], dtype=np.float32)
    want = (p.astype(np.float64) + np.array(c)).astype(np.float32)
    assert np.array_equal(dst.attribute("Position3D")[0], want)
    assert want[1] == np.float32(16777216.0)  # 16777217 is not an f32: the sum is formed in f64 and rounds to even
    dst = O.OBuffer(O.OLayout.from_attributes([("Position3D", O.VEC3I32)]), 1, True)
    with pytest.raises(O.OracleError):  # "Unsupported datatype"
        O.pnts_read_into(body, [("Position3D", O.VEC3F32, 0)], 0, 1, dst, c)


def test_write_default_layout_roundtrip():  # pnts_writer.rs:448-500
    layout = O.OLayout.from_attributes([("Position3D", O.VEC3F32), ("ColorRGBA", O.VEC4U8), ("ColorRGB", O.VEC3U8),
                                        ("Normal", O.VEC3F32)], packed=1)
    src = O.OBuffer(layout, 2, True)
    src.set_attribute("Position3D", [[1, 2, 3], [2, 4, 6]])
    src.columns[1][:8] = [11, 21, 31, 41, 22, 44, 66, 88]
    src.set_attribute("ColorRGB", [[10, 20, 30], [20, 40, 60]])
    src.set_attribute("Normal", [[0.1, 0.2, 0.3], [0.2, 0.4, 0.6]])
    attrs, body = O.pnts_feature_table_body(src)
    assert [(a[0], a[2]) for a in attrs] == [("Position3D", 0), ("ColorRGBA", 24), ("ColorRGB", 32), ("Normal", 40)]
    assert len(body) == 64 and body[38:40] == b"\0\0"  # 6 colour bytes padded to 8
    back = O.OBuffer(layout, 2, True)
    O.pnts_read_into(body, attrs, 0, 2, back)
    for i in range(4):
        assert np.array_equal(back.attribute_bytes(i), src.attribute_bytes(i))


def test_write_custom_layout():  # pnts_writer.rs:502-597: f64 positions -> f32, u16 colours -> u8 (`as`: low byte), intensity dropped
    layout = O.OLayout.from_attributes([("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16), ("Intensity", O.U16)], packed=1)
    src = O.OBuffer(layout, 2, False)
    src.set_attribute("Position3D", [[1, 2, 3], [2, 4, 6]])
    src.set_attribute("ColorRGB", [[0x1111, 0x2222, 0x3333], [0x2222, 0x4444, 0x6666]])
    src.set_attribute("Intensity", [10000, 20000])
    attrs, body = O.pnts_feature_table_body(src)
    assert [a[:2] for a in attrs] == [("Position3D", O.VEC3F32), ("ColorRGB", O.VEC3U8)]
    back = O.OBuffer(O.OLayout.from_attributes([a[:2] for a in attrs], packed=1), 2, True)
    O.pnts_read_into(body, attrs, 0, 2, back)
    assert back.attribute("Position3D").tolist() == [[1, 2, 3], [2, 4, 6]]
    assert back.attribute("ColorRGB").tolist() == [[0x11, 0x22, 0x33], [0x22, 0x44, 0x66]]
