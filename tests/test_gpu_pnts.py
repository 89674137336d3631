"""GPU parity of the .pnts point path (pasture-io/src/tiles3d): the reference's fixture file, its reader/writer tests,
and oracle comparisons over random data, layouts and windows."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import HashMapBuffer, VectorBuffer, tiles3d
from pasture_b200 import PointAttributeDataType as DT
from tests import util
from tests.pnts_expected import check_fixture_arrays, fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("buffer_type", [VectorBuffer, HashMapBuffer])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_reference_fixture(buffer_type, device):
    blob, exp = fixture()
    reader = tiles3d.PntsReader(blob)
    assert reader.get_metadata().number_of_points() == exp["points_length"]
    names = [(m.name(), int(m.datatype())) for m in reader.get_default_point_layout().attributes()]
    assert names == [("Position3D", O.VEC3F32), ("ColorRGB", O.VEC3U8)]
    points = reader.read(8000, buffer_type, device)
    check_fixture_arrays(points.view_attribute("Position3D"), points.view_attribute("ColorRGB"), exp)
    with pytest.raises(EOFError):  # "No points remaining in PNTS file"
        reader.read_into(points, 1)


def test_reference_fixture_windows_into_other_layout():
    blob, exp = fixture()
    reader = tiles3d.PntsReader(torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda())  # body stays in HBM
    ol, pl = util.layouts([("Intensity", O.U16), ("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16)])
    dst = VectorBuffer(pl, 100, "cuda")
    dst.set_attribute("Intensity", np.arange(100))
    reader.seek_point(7900)
    assert reader.read_into(dst, 1000) == 100
    assert dst.view_attribute("Position3D")[-4:].tolist() == exp["last_positions"]
    assert dst.view_attribute("ColorRGB")[-4:].tolist() == exp["last_rgb"]
    assert np.array_equal(dst.view_attribute("Intensity"), np.arange(100))


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_pnts_reader_read_modes(device):  # pnts_reader.rs:420-482
    layout = pb.PointLayout.from_attributes_packed([pb.attributes.POSITION_3D.with_custom_datatype(DT.Vec3f32)], 1)
    points = HashMapBuffer(layout, 2, device)
    points.set_attribute("Position3D", np.array([[10, 10, 10], [20, 20, 20]], dtype=np.float32))
    writer = tiles3d.PntsWriter(layout)
    writer.set_rtc_center([10.0, 10.0, 10.0])
    writer.write(points)
    image = writer.flush()
    reader = tiles3d.PntsReader(image)
    got = reader.read(reader.get_metadata().number_of_points(), VectorBuffer, device)
    assert got.view_attribute("Position3D").tolist() == [[20, 20, 20], [30, 30, 30]]
    reader = tiles3d.PntsReader(image)
    reader.set_read_positions_mode(tiles3d.RELATIVE_TO_CENTER)
    got = reader.read(2, VectorBuffer, device)
    assert got.view_attribute("Position3D").tolist() == [[10, 10, 10], [20, 20, 20]]


def test_write_pnts_default_layout():  # pnts_writer.rs:448-500
    attrs = [("Position3D", O.VEC3F32), ("ColorRGBA", O.VEC4U8), ("ColorRGB", O.VEC3U8), ("Normal", O.VEC3F32)]
    ol, pl = util.layouts(attrs, packed=1)
    src = O.OBuffer(ol, 2, True)
    src.set_attribute("Position3D", [[1, 2, 3], [2, 4, 6]])
    src.columns[1][:8] = [11, 21, 31, 41, 22, 44, 66, 88]
    src.set_attribute("ColorRGB", [[10, 20, 30], [20, 40, 60]])
    src.set_attribute("Normal", [[0.1, 0.2, 0.3], [0.2, 0.4, 0.6]])
    writer = tiles3d.PntsWriter(pl)
    writer.write(util.to_pb(src, pl, "cuda"))
    image = writer.flush()
    want_attrs, want_body = O.pnts_feature_table_body(src)
    reader = tiles3d.PntsReader(image)
    body0 = min(reader.attribute_offsets.values())
    assert image[body0:body0 + len(want_body)] == want_body
    assert len(image) % 8 == 0 and int.from_bytes(image[8:12], "little") == len(image)
    back = reader.read(2, HashMapBuffer, "cuda")
    assert back.point_layout() == pl
    util.assert_buffers_match(src, back)


def test_write_pnts_custom_layout():  # pnts_writer.rs:502-597
    ol, pl = util.layouts([("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16), ("Intensity", O.U16)], packed=1)
    src = O.OBuffer(ol, 2, False)
    src.set_attribute("Position3D", [[1, 2, 3], [2, 4, 6]])
    src.set_attribute("ColorRGB", [[0x1111, 0x2222, 0x3333], [0x2222, 0x4444, 0x6666]])
    src.set_attribute("Intensity", [10000, 20000])
    writer = tiles3d.PntsWriter(pl)
    writer.write(util.to_pb(src, pl, "cuda"))
    reader = tiles3d.PntsReader(writer.flush())
    got = reader.read(2, HashMapBuffer, "cuda")
    want_layout = pb.PointLayout.from_attributes_packed([pb.attributes.POSITION_3D.with_custom_datatype(DT.Vec3f32),
                                                         pb.attributes.COLOR_RGB.with_custom_datatype(DT.Vec3u8)], 1)
    assert got.point_layout() == want_layout
    assert got.view_attribute("Position3D").tolist() == [[1, 2, 3], [2, 4, 6]]
    assert got.view_attribute("ColorRGB").tolist() == [[0x11, 0x22, 0x33], [0x22, 0x44, 0x66]]


SRC_LAYOUTS = [
    [("Position3D", O.VEC3F64), ("Intensity", O.U16), ("ColorRGB", O.VEC3U16), ("Normal", O.VEC3F32), ("GpsTime", O.F64)],
    [("ColorRGBA", O.VEC4U8), ("Position3D", O.VEC3F32), ("Classification", O.U8)],
    [("Normal", O.VEC3F32), ("ColorRGB", O.VEC3U8)],
    [("Intensity", O.U16)],
]


@pytest.mark.parametrize("li", range(len(SRC_LAYOUTS)))
@pytest.mark.parametrize("columnar", [True, False])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
@pytest.mark.parametrize("n", [1, 7, 1000, 40001])
def test_write_body_matches_oracle(li, columnar, device, n):
    ol, pl = util.layouts(SRC_LAYOUTS[li], packed=1 if li % 2 else 0)
    ob, pbuf = util.random_bytes_buffers(ol, pl, n, columnar, seed=100 + n + li, device=device)
    want_attrs, want_body = O.pnts_feature_table_body(ob)
    writer = tiles3d.PntsWriter(pl)
    writer.write(pbuf)
    total, offsets, body = writer._feature_table_body()
    assert total == n and offsets == [a[2] for a in want_attrs]
    assert body == want_body
    assert [(d.name(), int(d.datatype())) for d in writer._attrs] == [a[:2] for a in want_attrs]


DST_LAYOUTS = [
    [("Position3D", O.VEC3F32), ("ColorRGB", O.VEC3U8)],
    [("GpsTime", O.F64), ("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16), ("Normal", O.VEC3F32)],
    [("ColorRGBA", O.VEC4U8), ("Intensity", O.U16)],
    [("Position3D", O.VEC3F64)],
]


@pytest.mark.parametrize("li", range(len(DST_LAYOUTS)))
@pytest.mark.parametrize("columnar", [True, False])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
@pytest.mark.parametrize("rtc", [None, [1234.5678, -0.001, 6378137.0]])
def test_read_matches_oracle(li, columnar, device, rtc):
    """a random 4-semantic file, read windows into differently typed / sized buffers, with and without RTC_CENTER"""
    n_file = 5000
    fl = [("Position3D", O.VEC3F32), ("ColorRGBA", O.VEC4U8), ("ColorRGB", O.VEC3U8), ("Normal", O.VEC3F32)]
    ol, pl = util.layouts(fl, packed=1)
    ob, pbuf = util.random_bytes_buffers(ol, pl, n_file, True, seed=77, device="cpu")
    writer = tiles3d.PntsWriter(pl)
    writer.write(pbuf)
    if rtc:
        writer.set_rtc_center(rtc)
    image = writer.flush()
    reader = tiles3d.PntsReader(image)
    file_attrs = [(m.name(), int(m.datatype()), reader.attribute_offsets[m.name()]) for m in reader.layout.attributes()]
    dol, dpl = util.layouts(DST_LAYOUTS[li])
    for first, count, buf_len in ((0, 5000, 5000), (123, 1000, 1000), (4990, 10, 64)):
        want, got = util.random_bytes_buffers(dol, dpl, buf_len, columnar, seed=first, device=device)
        if rtc and dol.index_by_name("Position3D") >= 0 and DST_LAYOUTS[li][dol.index_by_name("Position3D")][1] == O.VEC3F32:
            pass
        O.pnts_read_into(image, file_attrs, first, count, want, rtc)
        reader.seek_point(first)
        assert reader.read_into(got, count) == count
        util.assert_buffers_match(want, got, f"first={first}")


def test_rtc_on_unsupported_position_type():
    layout = pb.PointLayout.from_attributes_packed([pb.attributes.POSITION_3D.with_custom_datatype(DT.Vec3f32)], 1)
    points = HashMapBuffer(layout, 2, "cuda")
    writer = tiles3d.PntsWriter(layout)
    writer.set_rtc_center([1.0, 2.0, 3.0])
    writer.write(points)
    reader = tiles3d.PntsReader(writer.flush())
    bad = HashMapBuffer(pb.PointLayout.from_attributes([pb.attributes.POSITION_3D.with_custom_datatype(DT.Vec3i32)]), 2, "cuda")
    with pytest.raises(pb.PastureB200Error) as e:  # "Unsupported datatype {other} for POSITION_3D attribute"
        reader.read_into(bad, 2)
    assert e.value.code == -10


def test_full_size_roundtrip():
    """50 M points: f64 positions + u16 colours -> .pnts body (f32 / u8) -> f64 / u16 buffer with RTC; checked through
    the f32 rounding identity and the low-byte rule, on device"""
    n = 50_000_000
    _, pl = util.layouts([("Position3D", O.VEC3F64), ("ColorRGB", O.VEC3U16), ("Intensity", O.U16)])
    src = HashMapBuffer(pl, n, "cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    pos = (torch.rand((n, 3), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 2000.0
    col = torch.randint(0, 65536, (n, 3), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    src.columns[0][: 24 * n] = pos.view(torch.uint8).reshape(-1)
    src.columns[1][: 6 * n] = col.view(torch.uint8).reshape(-1)
    writer = tiles3d.PntsWriter(pl)
    writer.write(src)
    n_chunk, lay, body = writer._chunks[0]
    assert n_chunk == n and lay == [(0, 12), (12 * n, 3)]
    f32 = body[: 12 * n].view(torch.float32).reshape(n, 3)
    assert torch.equal(f32, pos.to(torch.float32))
    u8 = body[12 * n: 15 * n].reshape(n, 3)
    assert torch.equal(u8, (col.to(torch.int32) & 0xFF).to(torch.uint8))
    # read back through the ABI with an RTC centre into the wide layout
    from pasture_b200._lib import check, lib
    import ctypes as C
    arr = tiles3d._attr_array([("Position3D", DT.Vec3f32, 0), ("ColorRGB", DT.Vec3u8, 12 * n)])
    dst = HashMapBuffer(pl, n, "cuda")
    d = dst.desc()
    rtc = (C.c_double * 3)(1e6, -2e6, 0.5)
    check(lib().pb200_pnts_read_points(pb.get_context()._h, C.c_void_p(body.data_ptr()), body.numel(), arr, 2, 0, n, C.byref(d), rtc))
    got = dst.columns[0][: 24 * n].view(torch.float64).reshape(n, 3)
    want = f32.to(torch.float64) + torch.tensor([1e6, -2e6, 0.5], dtype=torch.float64, device="cuda")
    assert torch.equal(got, want)
    assert torch.equal(dst.columns[1][: 6 * n].view(torch.uint16).reshape(n, 3).to(torch.int32), u8.to(torch.int32))
