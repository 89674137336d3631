"""GPU parity tests of the conversion path (K0-K4), through the C ABI, against the CPU oracle.

Bit-exact everywhere (integer casts, byte copies and -- with FMA contraction off -- the f64 transforms).
Mirrors the reference's own tests: buffer_conversion.rs:684-930 (property tests for all four buffer pairs),
raw_readers.rs:670-1086 (LAS fixtures incl. the different-layout read)."""
import itertools

import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import attributes as A, PointAttributeDefinition, BufferLayoutConverter, VectorBuffer, HashMapBuffer
from pasture_b200 import PointAttributeDataType as DT
from tests import las_expected as E
from tests import util

pytestmark = pytest.mark.gpu

PAIRS = list(itertools.product([False, True], [False, True]))
BUF = {False: VectorBuffer, True: HashMapBuffer}
BIG = [("GpsTime", O.F64), ("ColorRGB", O.VEC3U16), ("Position3D", O.VEC3F64), ("Classification", O.U8), ("Intensity", O.I16)]
SMALL = [("Position3D", O.VEC3F64), ("Classification", O.U8)]


@pytest.fixture(scope="module")
def ctx():
    return pb.get_context(0)


def run_both(ocv, pcv, osrc, psrc, dst_col, ol_to, pl_to, device="cuda"):
    odst = ocv.convert(osrc, dst_col)
    pdst = pcv.convert(psrc, BUF[dst_col], device=device)
    torch.cuda.synchronize()
    util.assert_buffers_match(odst, pdst)
    return odst, pdst


# ---- the reference's property tests ----------------------------------------------------------------

@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_default(ctx, src_col, dst_col):  # buffer_conversion.rs:684-722
    ol, pl = util.layouts(BIG, packed=1)
    olt, plt = util.layouts(SMALL, packed=1)
    osrc, psrc = util.random_bytes_buffers(ol, pl, 16, src_col, seed=1)
    _, pdst = run_both(O.OConverter(ol, olt), BufferLayoutConverter.for_layouts(pl, plt), osrc, psrc, dst_col, olt, plt)
    assert np.array_equal(pdst.view_attribute(A.POSITION_3D), psrc.view_attribute(A.POSITION_3D))
    assert np.array_equal(pdst.view_attribute(A.CLASSIFICATION), psrc.view_attribute(A.CLASSIFICATION))


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_multiple_attributes_from_one(ctx, src_col, dst_col):  # :724-762
    ol, pl = util.layouts(BIG, packed=1)
    olt, plt = util.layouts([("Classification", O.U8), ("ReturnNumber", O.U8)])
    osrc, psrc = util.random_bytes_buffers(ol, pl, 16, src_col, seed=2)
    ocv = O.OConverter(ol, olt, with_default=True)
    ocv.set_custom_mapping(("Classification", O.U8), ("ReturnNumber", O.U8))
    pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
    pcv.set_custom_mapping(A.CLASSIFICATION, A.RETURN_NUMBER)
    _, pdst = run_both(ocv, pcv, osrc, psrc, dst_col, olt, plt)
    assert np.array_equal(pdst.view_attribute(A.RETURN_NUMBER), psrc.view_attribute(A.CLASSIFICATION))


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
@pytest.mark.parametrize("apply_to_source", [True, False])
def test_buffer_converter_transformed_attribute(ctx, src_col, dst_col, apply_to_source):  # :764-846
    ol, pl = util.layouts(BIG, packed=1)
    olt, plt = util.layouts([("Position3D", O.VEC3F64)])
    osrc, psrc = util.random_bytes_buffers(ol, pl, 16, src_col, seed=3)
    t = pb.Add(42.0)
    ocv = O.OConverter(ol, olt, with_default=True)
    ocv.set_custom_mapping_with_transformation(("Position3D", O.VEC3F64), ("Position3D", O.VEC3F64), O.VEC3F64,
                                               util.oracle_transform(t), apply_to_source)
    pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
    pcv.set_custom_mapping_with_transformation(A.POSITION_3D, A.POSITION_3D, t, apply_to_source)
    _, pdst = run_both(ocv, pcv, osrc, psrc, dst_col, olt, plt)
    assert np.array_equal(pdst.view_attribute(A.POSITION_3D), psrc.view_attribute(A.POSITION_3D) + 42.0)


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_buffer_converter_identity(ctx, src_col, dst_col):  # :848-873
    ol, pl = util.layouts(BIG, packed=1)
    osrc, psrc = util.random_bytes_buffers(ol, pl, 16, src_col, seed=4)
    run_both(O.OConverter(ol, ol, with_default=True), BufferLayoutConverter.for_layouts_with_default(pl, pl), osrc, psrc,
             dst_col, ol, pl)


def test_buffer_converter_mismatched_len(ctx):  # :912-930 should_panic
    ol, pl = util.layouts(BIG, packed=1)
    _, psrc = util.random_bytes_buffers(ol, pl, 16, False, seed=5)
    dst = VectorBuffer(pl, 8, "cuda")
    with pytest.raises(pb.PastureB200Error) as e:
        BufferLayoutConverter.for_layouts_with_default(pl, pl).convert_into(psrc, dst)
    assert e.value.code == -5


def test_contract_violations(ctx):
    ol, pl = util.layouts(BIG, packed=1)
    _, plt = util.layouts([("ReturnNumber", O.U8)])
    with pytest.raises(pb.PastureB200Error) as e:  # :112-116
        BufferLayoutConverter.for_layouts(pl, plt)
    assert e.value.code == -1
    assert BufferLayoutConverter.for_layouts_with_default(pl, plt).num_mappings() == 0
    _, a = util.layouts([("X", O.VEC4U8)])
    _, b = util.layouts([("X", O.U32)])
    with pytest.raises(pb.PastureB200Error) as e:  # :383-388
        BufferLayoutConverter.for_layouts(a, b)
    assert e.value.code == -2
    _, p32 = util.layouts([("Position3D", O.VEC3F32)])
    cv = BufferLayoutConverter.for_layouts_with_default(pl, p32)
    with pytest.raises(pb.PastureB200Error) as e:  # :209-213 transform after the cast must be typed as the target
        cv.set_custom_mapping_with_transformation(A.POSITION_3D, A.POSITION_3D.with_custom_datatype(DT.Vec3f32),
                                                  pb.Add(1.0, datatype=DT.Vec3f64), False)
    assert e.value.code == -3
    with pytest.raises(pb.PastureB200Error) as e:  # layout mismatch :302-303
        _, other = util.layouts(BIG)
        BufferLayoutConverter.for_layouts_with_default(pl, pl).convert_into(VectorBuffer(pl, 4, "cuda"), VectorBuffer(other, 4, "cuda"))
    assert e.value.code == -4


# ---- LAS fixtures (golden vectors of the reference) ---------------------------------------------------

@pytest.mark.parametrize("fmt", range(11))
@pytest.mark.parametrize("kind", ["plain", "extra_bytes"])
@pytest.mark.parametrize("columnar", [False, True])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_las_fixture_read_default_layout(ctx, las_fixtures, fmt, kind, columnar, device):
    entry = las_fixtures[kind][str(fmt)]
    ol_raw, pl_raw = util.las_layouts(fmt, True)
    if kind == "extra_bytes":
        ol_raw.add_attribute("ExtraBytesU32", O.U32, packed=1)
        pl_raw.add_attribute(PointAttributeDefinition("ExtraBytesU32", DT.U32), pb.FieldAlignment.Packed(1))
    assert pl_raw.size_of_point_entry() == entry["record_length"]
    raw = np.frombuffer(bytes.fromhex(entry["records_hex"]), dtype=np.uint8)
    src = VectorBuffer.from_bytes(pl_raw, raw, device)
    _, pl_def = util.las_layouts(fmt, False)
    cv = pb.get_default_las_converter(pl_raw, pl_def, entry["scale"], entry["offset"])
    dst = cv.convert(src, BUF[columnar])
    torch.cuda.synchronize()
    for name, expect in E.expected_default_layout_values(fmt).items():
        got = dst.view_attribute(name)
        assert got.dtype == expect.dtype and np.array_equal(got, expect), (name, got, expect)


@pytest.mark.parametrize("fmt", range(11))
@pytest.mark.parametrize("columnar", [False, True])
def test_las_fixture_read_different_layout(ctx, las_fixtures, fmt, columnar):  # raw_readers.rs:815-911
    entry = las_fixtures["plain"][str(fmt)]
    _, pl_raw = util.las_layouts(fmt, True)
    src = VectorBuffer.from_bytes(pl_raw, np.frombuffer(bytes.fromhex(entry["records_hex"]), dtype=np.uint8), "cuda")
    target = pb.PointLayout.from_attributes([
        A.POSITION_3D.with_custom_datatype(DT.Vec3f32), A.CLASSIFICATION.with_custom_datatype(DT.U32),
        A.COLOR_RGB.with_custom_datatype(DT.Vec3u8), A.POINT_SOURCE_ID, A.WAVEFORM_PARAMETERS])
    cv = pb.get_default_las_converter(pl_raw, target, entry["scale"], entry["offset"])
    dst = cv.convert(src, BUF[columnar])
    fl = E.fmt_flags(fmt)
    assert np.array_equal(dst.view_attribute("Position3D"), E.POSITIONS.astype(np.float32))
    assert np.array_equal(dst.view_attribute("Classification"), E.CLASSIFICATIONS.astype(np.uint32))
    assert np.array_equal(dst.view_attribute("ColorRGB"), E.COLORS.astype(np.uint8) if fl["color"] else np.zeros((10, 3), np.uint8))
    assert np.array_equal(dst.view_attribute("PointSourceID"), E.POINT_SOURCE_IDS)
    assert np.array_equal(dst.view_attribute("WaveformParameters"),
                          E.WAVEPACKET_PARAMETERS if fl["waveform"] else np.zeros((10, 3), np.float32))


def test_las_invalid_position_type(ctx):  # raw_readers.rs:56
    _, raw = util.las_layouts(0, True)
    _, t = util.layouts([("Position3D", O.VEC3I32)])
    with pytest.raises(pb.PastureB200Error) as e:
        pb.get_default_las_converter(raw, t, (1, 1, 1), (0, 0, 0))
    assert e.value.code == -10


# ---- every cast of the table, incl. the Rust `as` edge cases ---------------------------------------

EDGE_F = [0.0, -0.0, 0.5, -0.5, 0.99, -0.99, 1.5, -1.5, 2.9, -2.9, 127.0, 127.9, 128.0, -128.0, -128.9, -129.0, 255.0, 255.9,
          256.0, 300.7, 32767.5, 32768.0, -32768.9, 65535.9, 65536.0, 2147483647.0, 2147483648.0, -2147483648.0,
          -2147483649.0, 4294967295.0, 4294967296.0, 9.223372036854775e18, 9.223372036854776e18, -9.223372036854776e18,
          -9.3e18, 1.8446744073709550e19, 1.8446744073709552e19, 1e20, -1e20, 1e40, -1e40, 1e-50, 16777217.0, 3.4028235e38,
          3.5e38, float("inf"), float("-inf"), float("nan"), 5e-324, 1.1754942e-38, 0.1, 1 / 3]


def edge_values(comp, n):
    rng = np.random.default_rng(np.dtype(comp).num + 100)
    if np.issubdtype(comp, np.floating):
        with np.errstate(over="ignore"):
            base = np.array(EDGE_F, dtype=np.float64).astype(comp)
        rnd = rng.integers(0, 256, (n - len(base), np.dtype(comp).itemsize), dtype=np.uint8).view(comp).reshape(-1)
        return np.concatenate([base, rnd])
    info = np.iinfo(comp)
    bits = 8 * np.dtype(comp).itemsize
    unsigned = np.dtype(f"u{np.dtype(comp).itemsize}")
    raw = [0, 1, int(info.max), int(info.min), int(info.max) - 1, int(info.min) + 1, 2, 127, 128, 255, 256, 16777217, -1, -2]
    base = np.array([v % (1 << bits) for v in raw], dtype=unsigned).view(comp)
    rnd = rng.integers(info.min, info.max, n - len(base), dtype=comp, endpoint=True)
    return np.concatenate([base, rnd])


@pytest.mark.parametrize("src_col,dst_col", [(True, True), (False, False)])
def test_all_scalar_casts_bit_exact(ctx, src_col, dst_col):
    """all 90 directed scalar pairs (attribute_conversion.rs:194-246) in one layout pair per source type"""
    n = 1000
    for sdt in range(10):
        src_attrs = [("v", sdt)]
        dst_types = [d for d in range(10) if d != sdt]
        ol, pl = util.layouts(src_attrs, packed=1)
        olt, plt = util.layouts([(f"t{d}", d) for d in dst_types], packed=1)
        osrc = O.OBuffer(ol, n, src_col)
        osrc.set_attribute("v", edge_values(O.NP_DTYPES[sdt], n))
        psrc = util.to_pb(osrc, pl)
        ocv = O.OConverter(ol, olt, with_default=True)
        pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
        for d in dst_types:
            ocv.set_custom_mapping(("v", sdt), (f"t{d}", d))
            pcv.set_custom_mapping(PointAttributeDefinition("v", sdt), PointAttributeDefinition(f"t{d}", d))
        run_both(ocv, pcv, osrc, psrc, dst_col, olt, plt)


def test_all_vec3_casts_bit_exact(ctx):
    """all 20 directed Vec3 pairs (attribute_conversion.rs:248-260)"""
    n = 999
    vec3 = [O.VEC3U8, O.VEC3U16, O.VEC3F32, O.VEC3I32, O.VEC3F64]
    for sdt in vec3:
        ol, pl = util.layouts([("v", sdt)], packed=1)
        dts = [d for d in vec3 if d != sdt]
        olt, plt = util.layouts([(f"t{d}", d) for d in dts], packed=1)
        for src_col, dst_col in PAIRS:
            osrc = O.OBuffer(ol, n, src_col)
            osrc.set_attribute("v", edge_values(O.NP_DTYPES[O.VEC3_COMPONENT[sdt]], 3 * n).reshape(n, 3))
            psrc = util.to_pb(osrc, pl)
            ocv = O.OConverter(ol, olt, with_default=True)
            pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
            for d in dts:
                ocv.set_custom_mapping(("v", sdt), (f"t{d}", d))
                pcv.set_custom_mapping(PointAttributeDefinition("v", sdt), PointAttributeDefinition(f"t{d}", d))
            run_both(ocv, pcv, osrc, psrc, dst_col, olt, plt)


# ---- random layouts, ranges, ragged sizes, both kernels ---------------------------------------------

@pytest.mark.parametrize("seed", range(12))
def test_random_layout_conversion(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    n_src = int(rng.integers(1, 9))
    src_attrs = [(f"a{i}", int(rng.choice(list(range(16)) + [O.BYTEARRAY])), int(rng.integers(1, 9))) for i in range(n_src)]
    src_attrs = [(a, d, e if d == O.BYTEARRAY else 0) for a, d, e in src_attrs]
    ol, pl = util.layouts(src_attrs, packed=int(rng.choice([0, 1, 2])))
    dst_attrs = []
    for (a, d, e) in src_attrs:
        if rng.random() < 0.25:
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
    order = rng.permutation(len(dst_attrs))
    dst_attrs = [dst_attrs[i] for i in order]
    olt, plt = util.layouts(dst_attrs, packed=int(rng.choice([0, 1, 4])))
    n = int(rng.choice([1, 7, 100, 1000, 4097, 20011]))
    for src_col, dst_col in PAIRS:
        osrc, psrc = util.random_bytes_buffers(ol, pl, n, src_col, seed=seed, finite_floats=bool(seed % 2))
        ocv = O.OConverter(ol, olt, with_default=True)
        pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
        # target pre-filled with a pattern: unmapped bytes / padding must survive
        odst = O.OBuffer(olt, n, dst_col)
        pdst = BUF[dst_col](plt, n, "cuda")
        if dst_col:
            for c in odst.columns:
                c[:] = 0xA5
            for c in pdst.columns:
                c.fill_(0xA5)
        else:
            odst.aos[:] = 0xA5
            pdst.data.fill_(0xA5)
        ocv.convert_into(osrc, odst)
        pcv.convert_into(psrc, pdst)
        torch.cuda.synchronize()
        util.assert_buffers_match(odst, pdst, f"seed {seed} {src_col}->{dst_col}")
        if not dst_col:
            assert np.array_equal(odst.aos[: n * olt.size], pdst.raw_bytes()), "padding / unmapped bytes changed"


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_fresh_target_conversion_writes_every_byte(ctx, seed, device):
    """`convert` semantics on caller-owned memory (buffer_conversion.rs:242-259: a zero-filled new buffer, then
    convert_into): pb200_converter_convert_fresh_range must leave the target range exactly like the oracle's `convert`,
    although it starts from garbage -- mapped attributes converted, unmapped attributes / padding zero, and the points
    outside the range untouched.  Random layouts with holes (missing attributes, alignment padding), all buffer pairs."""
    import ctypes as C
    from pasture_b200._lib import check, lib
    rng = np.random.default_rng(4000 + seed)
    n_src = int(rng.integers(2, 8))
    src_attrs = [(f"a{i}", int(rng.choice(list(range(16)))), 0) for i in range(n_src)]
    ol, pl = util.layouts(src_attrs, packed=int(rng.choice([0, 1])))
    dst_attrs = [(a, d, e) for (a, d, e) in src_attrs if rng.random() < 0.7] or [src_attrs[0]]
    dst_attrs.insert(int(rng.integers(0, len(dst_attrs) + 1)), ("hole_u8", O.U8, 0))      # no source: stays default
    dst_attrs.insert(int(rng.integers(0, len(dst_attrs) + 1)), ("hole_f64", O.F64, 0))
    olt, plt = util.layouts(dst_attrs, packed=int(rng.choice([0, 1, 4])))
    n, lo, hi = 3001, 5, 2990
    for src_col, dst_col in PAIRS:
        osrc, psrc = util.random_bytes_buffers(ol, pl, n, src_col, seed=seed, finite_floats=True)
        if device == "cpu":
            psrc = psrc.to("cpu")
        ocv = O.OConverter(ol, olt, with_default=True)
        pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
        want = ocv.convert(osrc, dst_col)  # zero-filled target, convert_into
        pdst = BUF[dst_col](plt, n, device)
        for t in (pdst.columns if dst_col else [pdst.data]):
            t.fill_(0xC3)
        sd, dd = psrc.desc(), pdst.desc()
        check(lib().pb200_converter_convert_fresh_range(pcv._h, C.byref(sd), lo, hi, C.byref(dd), lo, hi, None))
        torch.cuda.synchronize()
        for i in range(len(plt)):
            got, exp = pdst._attribute_bytes(i), want.attribute_bytes(i)
            assert np.array_equal(got[lo:hi], exp[lo:hi]), (seed, src_col, dst_col, plt.at(i))
            assert np.all(got[:lo] == 0xC3) and np.all(got[hi:] == 0xC3), "points outside the range were touched"
        if not dst_col:  # padding bytes inside the range are zero as well
            rec = plt.size_of_point_entry()
            raw = pdst.raw_bytes().reshape(n, rec)
            assert np.array_equal(raw[lo:hi], want.aos[: n * rec].reshape(n, rec)[lo:hi])
        # and the allocating convenience call: uninitialised target, whole range
        full = pcv.convert(psrc, BUF[dst_col], device=device)
        util.assert_buffers_match(want, full)


@pytest.mark.parametrize("src_col,dst_col", PAIRS)
def test_ranges_with_unaligned_offsets(ctx, src_col, dst_col):
    """convert_into_range (buffer_conversion.rs:292): arbitrary sub-ranges, neighbours untouched"""
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    n = 5000
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[:] = O.gen_las_fmt0_records(0, n)
    if src_col:
        osrc = O.OConverter(ol, ol, with_default=True).convert(osrc, True)
    psrc = util.to_pb(osrc, pl)
    ocv = O.OConverter.las_default(ol, olt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
    pcv = pb.get_default_las_converter(pl, plt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
    odst = O.OBuffer(olt, n + 13, dst_col)
    pdst = BUF[dst_col](plt, n + 13, "cuda")
    for (sb, se, db) in [(0, 1, 0), (3, 1003, 7), (1001, 4998, 1010), (4999, 5000, 5012), (17, 17, 3)]:
        ocv.convert_into_range(osrc, sb, se, odst, db, db + (se - sb))
        pcv.convert_into_range(psrc, range(sb, se), pdst, range(db, db + (se - sb)))
    torch.cuda.synchronize()
    util.assert_buffers_match(odst, pdst)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 767, 768, 769, 100003])
def test_c2_sizes_tile_and_direct_kernels_agree(ctx, n):
    """C2 mapping at ragged sizes; the tile pipeline (K1) and the direct kernel (K0) give identical bytes"""
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    osrc = O.OBuffer(ol, n, False)
    if n:
        osrc.aos[: 20 * n] = O.gen_las_fmt0_records(0, n)
    psrc = util.to_pb(osrc, pl)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    odst = O.OConverter.las_default(ol, olt, scale, offset).convert(osrc, True)
    pcv = pb.get_default_las_converter(pl, plt, scale, offset)
    a = pcv.convert(psrc, HashMapBuffer)
    ctx.set_param("convert.force_direct", 1)
    try:
        b = pcv.convert(psrc, HashMapBuffer)
    finally:
        ctx.set_param("convert.force_direct", 0)
    torch.cuda.synchronize()
    util.assert_buffers_match(odst, a, "tile kernel")
    util.assert_buffers_match(odst, b, "direct kernel")


@pytest.mark.parametrize("tile,threads,stages,cps", [(16, 32, 1, 1), (128, 64, 2, 4), (2048, 512, 4, 1), (640, 256, 3, 2)])
def test_pipeline_shapes(ctx, tile, threads, stages, cps):
    """every tile/stage/thread shape of the pipeline produces the same bytes"""
    ol, pl = util.las_layouts(3, True)
    olt, plt = util.las_layouts(3, False)
    n = 30011
    osrc, psrc = util.random_bytes_buffers(ol, pl, n, False, seed=9)
    odst = O.OConverter.las_default(ol, olt, (0.01, 0.02, 0.03), (1.5, -2.5, 3.5)).convert(osrc, False)
    pcv = pb.get_default_las_converter(pl, plt, (0.01, 0.02, 0.03), (1.5, -2.5, 3.5))
    for k, v in (("tile_points", tile), ("threads", threads), ("stages", stages), ("ctas_per_sm", cps)):
        ctx.set_param("convert." + k, v)
    try:
        pdst = pcv.convert(psrc, VectorBuffer)
        torch.cuda.synchronize()
    finally:
        for k in ("tile_points", "threads", "stages", "ctas_per_sm"):
            ctx.set_param("convert." + k, 0)
    util.assert_buffers_match(odst, pdst)


def test_write_direction_c1_and_out_of_range_count(ctx):
    """C1: LasPointFormat0 (35 B) -> raw LAS fmt-0 (20 B) with (p - offset) / scale before the cast
    (write_helpers.rs:15-17: truncation; out-of-range values counted instead of panicking)"""
    n = 200000
    offset = (500000.0, 5400000.0, 100.0)
    ol, pl = util.las_layouts(0, False)
    olt, plt = util.las_layouts(0, True)
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[:] = O.gen_c1_points(0, n, 42, offset)
    pos = osrc.attribute("Position3D").copy()
    pos[5] = [3e6 + offset[0], 0, 0]       # > i32::MAX mm
    pos[6] = [np.nan, np.inf, -np.inf]     # NaN -> 0 (no panic), +-inf out of range
    osrc.set_attribute("Position3D", pos)
    psrc = util.to_pb(osrc, pl)
    t = pb.InvScaleOffset(0.001, offset)
    ocv = O.OConverter(ol, olt, with_default=True)
    ocv.set_custom_mapping_with_transformation(("Position3D", O.VEC3F64), ("LASLocalPosition", O.VEC3I32), O.VEC3F64,
                                               util.oracle_transform(t), True)
    pcv = BufferLayoutConverter.for_layouts_with_default(pl, plt)
    pcv.set_custom_mapping_with_transformation(A.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, t, True)
    odst = ocv.convert(osrc, False)
    pdst = VectorBuffer(plt, n, "cuda")
    oor = pcv.convert_into(psrc, pdst, count_out_of_range=True)
    util.assert_buffers_match(odst, pdst)
    assert oor == 4  # x, y of point 5 (y = -5.4e9 mm) and +-inf of point 6
    got = pdst.view_attribute(pb.ATTRIBUTE_LOCAL_LAS_POSITION)
    exact = np.trunc((pos[100:] - np.array(offset)) / 0.001)
    assert np.array_equal(got[100:], exact.astype(np.int32))


def test_host_memspace_staged_path(ctx):
    """HOST buffers: chunks staged through device memory; result lands in caller-owned host memory"""
    ol, pl = util.las_layouts(1, True)
    olt, plt = util.las_layouts(1, False)
    n = 70001
    osrc, _ = util.random_bytes_buffers(ol, pl, n, False, seed=21, device="cpu")
    psrc = util.to_pb(osrc, pl, "cpu", pinned=True)
    for dst_col in (False, True):
        odst = O.OConverter.las_default(ol, olt, (0.001,) * 3, (1.0, 2.0, 3.0)).convert(osrc, dst_col)
        pcv = pb.get_default_las_converter(pl, plt, (0.001,) * 3, (1.0, 2.0, 3.0))
        pdst = BUF[dst_col](plt, n, "cpu", pinned=True)
        pcv.convert_into(psrc, pdst)
        util.assert_buffers_match(odst, pdst)
        # mixed: host source, device target
        pdev = BUF[dst_col](plt, n, "cuda")
        pcv.convert_into(psrc, pdev)
        torch.cuda.synchronize()
        util.assert_buffers_match(odst, pdev)


def test_fused_bounds_equals_oracle_bounds(ctx):
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    n = 123457
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[:] = O.gen_las_fmt0_records(0, n)
    psrc = util.to_pb(osrc, pl)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    odst = O.OConverter.las_default(ol, olt, scale, offset).convert(osrc, True)
    omn, omx = O.calculate_bounds(odst)
    pcv = pb.get_default_las_converter(pl, plt, scale, offset)
    pdst = HashMapBuffer(plt, n, "cuda")
    mn, mx = pcv.convert_into_range_with_bounds(psrc, range(0, n), pdst, range(0, n))
    assert list(mn) == list(omn) and list(mx) == list(omx)
    util.assert_buffers_match(odst, pdst)
    assert pcv.convert_into_range_with_bounds(psrc, range(0, 0), pdst, range(0, 0)) is None


def test_transform_attribute_and_converting_view(ctx):
    ol, pl = util.layouts([("Position3D", O.VEC3F32), ("Intensity", O.U16), ("Position3D64", O.VEC3F64)], packed=1)
    n = 3001
    for col in (False, True):
        osrc, psrc = util.random_bytes_buffers(ol, pl, n, col, seed=33)
        # pnts_reader.rs:265-277 RTC centre on Vec3f32 and Vec3f64
        rtc = (1215019.0, -4736339.0, 4081627.0)
        pb.transform_attribute(psrc, A.POSITION_3D.with_custom_datatype(DT.Vec3f32), pb.Add(rtc))
        pb.transform_attribute(psrc, PointAttributeDefinition("Position3D64", DT.Vec3f64), pb.Add(rtc))
        torch.cuda.synchronize()
        p32 = osrc.attribute("Position3D")
        assert np.array_equal(psrc.view_attribute("Position3D").view(np.uint32),
                              (p32.astype(np.float64) + np.array(rtc)).astype(np.float32).view(np.uint32))
        assert np.array_equal(psrc.view_attribute("Position3D64").view(np.uint64),
                              (osrc.attribute("Position3D64") + np.array(rtc)).view(np.uint64))
        assert np.array_equal(psrc.view_attribute("Intensity"), osrc.attribute("Intensity"))
        # AttributeViewConverting: intensity as f64, positions as Vec3i32
        v = pb.view_attribute_with_conversion(psrc, A.INTENSITY.with_custom_datatype(DT.F64))
        assert np.array_equal(v, osrc.attribute("Intensity").astype(np.float64))
        with pytest.raises(pb.PastureB200Error):
            pb.view_attribute_with_conversion(psrc, A.INTENSITY.with_custom_datatype(DT.Vec3f64))


@pytest.mark.parametrize("columnar,device", [(False, "cuda"), (True, "cuda"), (True, "cpu")])
def test_converting_views_and_transform_attribute_against_the_oracle(ctx, columnar, device):
    """Row W against the ORACLE (not numpy): AttributeViewConverting (buffer_views.rs:533-650) materialises an attribute
    through the same cast table as the converter -- checked for every scalar and Vec3 view type of arbitrary-bit-pattern
    sources (NaN, inf, saturation) via the oracle's single-mapping conversion -- and transform_attribute
    (point_buffer.rs:391-404) equals the oracle's same-layout conversion with the transform on the attribute"""
    attrs = [("s_f64", O.F64), ("s_i16", O.I16), ("s_u64", O.U64), ("v_f32", O.VEC3F32), ("v_i32", O.VEC3I32), ("Position3D", O.VEC3F64)]
    ol, pl = util.layouts(attrs, packed=1)
    n = 4099
    osrc, _ = util.random_bytes_buffers(ol, pl, n, columnar, seed=91, device=device, finite_floats=True)
    special = np.array([np.inf, -np.inf, np.nan, 1e300, -1e300, 3e9, -3e9, 70000.5, -0.0, 255.9999])
    for name in ("s_f64", "v_f32", "Position3D"):  # non-finite / saturating / truncating inputs in known places
        v = osrc.attribute(name).copy()
        flat = v.reshape(n, -1)
        with np.errstate(over="ignore"):
            flat[10:10 + len(special), 0] = special.astype(flat.dtype)
        osrc.set_attribute(name, v)
    psrc = util.to_pb(osrc, pl, device)

    def same(got, want):  # bit-equal, except that a NaN may carry any payload (Rust's `as` does not specify it either)
        got, want = np.ascontiguousarray(got), np.ascontiguousarray(want)
        if np.issubdtype(want.dtype, np.floating):
            return np.array_equal(got, want, equal_nan=True) and np.array_equal(np.signbit(got), np.signbit(want))
        return np.array_equal(got, want)
    for name, dtype, _ in [(a, d, 0) for a, d in attrs]:
        views = range(10) if dtype <= O.F64 else range(O.VEC3U8, O.VEC3F64 + 1)
        for vd in views:
            if vd == dtype:
                continue
            olt, _ = util.layouts([(name, int(vd))])
            want = O.OConverter(ol, olt).convert(osrc, True).attribute(name)
            got = pb.view_attribute_with_conversion(psrc, PointAttributeDefinition(name, int(vd)))
            assert same(got, want), (name, dtype, vd)
    # transform_attribute: v * s + o on the Vec3f64 positions, v + c on the f64 scalar, in place; everything else untouched
    t1, t2 = pb.ScaleOffset((0.01, 0.02, 0.03), (5.0, -6.0, 7.0)), pb.Add((1234.5, 0.0, 0.0))
    ocv = O.OConverter(ol, ol, with_default=True)
    ocv.set_custom_mapping_with_transformation(("Position3D", O.VEC3F64), ("Position3D", O.VEC3F64), O.VEC3F64, util.oracle_transform(t1), True)
    ocv.set_custom_mapping_with_transformation(("s_f64", O.F64), ("s_f64", O.F64), O.F64, util.oracle_transform(t2), True)
    want = ocv.convert(osrc, columnar)
    pb.transform_attribute(psrc, A.POSITION_3D, t1)
    pb.transform_attribute(psrc, PointAttributeDefinition("s_f64", DT.F64), t2)
    torch.cuda.synchronize()
    for name, _ in attrs:
        assert same(psrc.view_attribute(name), want.attribute(name)), name


def test_large_range_matches_direct_kernel_and_torch(ctx):
    """6 M points (many tiles per CTA): tile pipeline == direct kernel == a torch recomputation of the positions"""
    n = 6_000_011
    pl, plt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    src = pb.algorithms.synth_las_fmt0_records(n)
    scale, offset = (0.001, 0.001, 0.001), (500000.0, 5400000.0, 100.0)
    cv = pb.get_default_las_converter(pl, plt, scale, offset)
    a = HashMapBuffer(plt, n, "cuda")
    bounds = cv.convert_into_range_with_bounds(src, range(0, n), a, range(0, n))
    a2 = HashMapBuffer(plt, n, "cuda")
    cv.convert_into(src, a2)                                                    
    cv.convert_into(src, a2)                                                    
    ctx.set_param("convert.force_direct", 1)
    try:
        b = cv.convert(src, HashMapBuffer)
    finally:
        ctx.set_param("convert.force_direct", 0)
    torch.cuda.synchronize()
    for i in range(len(plt)):
        assert torch.equal(a.columns[i], b.columns[i]) and torch.equal(a2.columns[i], b.columns[i]), plt.at(i)
    xyz = src.data[: 20 * n].view(n, 20)[:, :12].contiguous().view(torch.int32).view(n, 3)
    pos = xyz.double() * 0.001 + torch.tensor(offset, dtype=torch.float64, device="cuda")  # mul then add: two roundings
    got = a.columns[0][: 24 * n].view(torch.float64).view(n, 3)
    assert torch.equal(got, pos)
    assert list(bounds[0]) == pos.min(0).values.tolist() and list(bounds[1]) == pos.max(0).values.tolist()


def test_c2_full_size_bit_exact_against_torch_and_direct_kernel(ctx):
    """BASELINE config C2 at its full 100 M points: every output column of the tile pipeline equals (a) the direct
    kernel (second implementation) and (b) an independent torch recomputation from the raw record bytes -- positions
    cast i32 -> f64, * scale, + offset as two roundings (raw_readers.rs:42-48), bit fields (:61-103), plain copies --
    and the write direction brings every coordinate back to within one LAS unit (truncation, write_helpers.rs:15-17)."""
    n = 100_000_000 if torch.cuda.get_device_properties(0).total_memory > 100e9 else 4_000_000
    pl, plt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    src = pb.algorithms.synth_las_fmt0_records(n)
    scale, offset = (0.001, 0.001, 0.001), (500000.0, 5400000.0, 100.0)
    cv = pb.get_default_las_converter(pl, plt, scale, offset)
    a = HashMapBuffer(plt, n, "cuda")
    bounds = cv.convert_into_range_with_bounds(src, range(0, n), a, range(0, n))
    ctx.set_param("convert.force_direct", 1)
    try:
        b = cv.convert(src, HashMapBuffer)
    finally:
        ctx.set_param("convert.force_direct", 0)
    for i in range(len(plt)):
        assert torch.equal(a.columns[i], b.columns[i]), plt.at(i)
    del b
    rec = src.data[: 20 * n].view(n, 20)
    xyz = rec[:, :12].contiguous().view(torch.int32).view(n, 3)
    pos = xyz.double() * 0.001 + torch.tensor(offset, dtype=torch.float64, device="cuda")
    assert torch.equal(a.columns[0][: 24 * n].view(torch.float64).view(n, 3), pos)
    assert list(bounds[0]) == pos.min(0).values.tolist() and list(bounds[1]) == pos.max(0).values.tolist()
    del pos
    names = [m.name() for m in plt.attributes()]
    col = lambda name: a.columns[names.index(name)]
    flags = rec[:, 14]
    assert torch.equal(col("Intensity")[: 2 * n].view(n, 2), rec[:, 12:14])
    assert torch.equal(col("ReturnNumber")[:n], flags & 7)
    assert torch.equal(col("NumberOfReturns")[:n], (flags >> 3) & 7)
    assert torch.equal(col("ScanDirectionFlag")[:n], (flags >> 6) & 1)
    assert torch.equal(col("EdgeOfFlightLine")[:n], (flags >> 7) & 1)
    assert torch.equal(col("Classification")[:n], rec[:, 15])
    assert torch.equal(col("ScanAngleRank")[:n], rec[:, 16])
    assert torch.equal(col("UserData")[:n], rec[:, 17])
    assert torch.equal(col("PointSourceID")[: 2 * n].view(n, 2), rec[:, 18:20])
    # write direction: (p - offset) / scale truncated back to i32
    back = pb.VectorBuffer(pl, n, "cuda")
    wr = pb.BufferLayoutConverter.for_layouts_with_default(plt, pl)
    wr.set_custom_mapping_with_transformation(A.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, pb.InvScaleOffset(0.001, offset), True)
    oor = wr.convert_into(a, back, count_out_of_range=True)
    xyz2 = back.data[: 20 * n].view(n, 20)[:, :12].contiguous().view(torch.int32).view(n, 3)
    assert int((xyz2 - xyz).abs().max()) <= 1
    assert oor == 0
    assert torch.equal(back.data[: 20 * n].view(n, 20)[:, 12:14], rec[:, 12:14])  # intensity survives the round trip


# ---- columnar -> packed interleaved records: the grouped copy (G consecutive records per lane) -----------------

GROUPED_LAYOUTS = {
    # record stride 35 (odd, G = 4): the default LAS format-0 layout, every copy size 1 / 2 / 24
    "las_default_0": None,
    # stride 41 (odd): 8-, 12-, 6-, 4-, 3-byte elements at every alignment residue
    "odd_41": [("p", O.VEC3F32, 0), ("t", O.F64, 0), ("c", O.VEC3U16, 0), ("i", O.U32, 0), ("rgb", O.VEC3U8, 0), ("k", O.U64, 0)],
    # stride 26 (= 2 mod 4, G = 2)
    "even_26": [("p", O.VEC3I32, 0), ("i", O.U16, 0), ("f", O.U8, 0), ("c", O.U8, 0), ("a", O.I8, 0), ("u", O.U8, 0), ("s", O.U16, 0), ("c3", O.VEC3U16, 0)],
    # stride 25 (odd) with a Vec3f64
    "odd_25": [("Position3D", O.VEC3F64, 0), ("cls", O.U8, 0)],
}


@pytest.mark.parametrize("name", list(GROUPED_LAYOUTS))
@pytest.mark.parametrize("n", [1, 127, 128, 129, 2048, 2049, 12345, 70001])
def test_columnar_to_packed_records_grouped_copy(ctx, name, n):
    """buffer_conversion.rs:494-544 (columnar -> interleaved) into packed records whose stride is not a multiple of 4: the
    tile kernel stores G consecutive records per lane (copy_loop_grouped).  Same bytes as the oracle, as the
    one-lane-per-record loops (convert.no_grouped_copy) and as the direct kernel, for whole buffers and for ranges whose
    first point makes the column tiles start at unaligned addresses; neighbours of the target range untouched."""
    if GROUPED_LAYOUTS[name] is None:
        ol, pl = util.las_layouts(0, False)
    else:
        ol, pl = util.layouts(GROUPED_LAYOUTS[name], packed=1)
    assert ol.size % 4 != 0
    osrc, psrc = util.random_bytes_buffers(ol, pl, n + 5, True, seed=n, finite_floats=False)
    ocv = O.OConverter(ol, ol, with_default=True)
    pcv = BufferLayoutConverter.for_layouts_with_default(pl, pl)
    for (sb, db) in [(0, 0), (1, 3), (5, 0)]:
        m = n + 5 - sb if sb else n
        results = []
        for param in (None, "no_grouped_copy", "force_direct"):
            odst = O.OBuffer(ol, m + 7, False)
            odst.aos[:] = 0x5A
            pdst = VectorBuffer(pl, m + 7, "cuda")
            pdst.data.fill_(0x5A)
            ocv.convert_into_range(osrc, sb, sb + m, odst, db, db + m)
            if param:
                ctx.set_param("convert." + param, 1)
            try:
                pcv.convert_into_range(psrc, range(sb, sb + m), pdst, range(db, db + m))
                torch.cuda.synchronize()
            finally:
                if param:
                    ctx.set_param("convert." + param, 0)
            assert np.array_equal(odst.aos[: (m + 7) * ol.size], pdst.raw_bytes()), (name, n, sb, db, param)


@pytest.mark.parametrize("dst", [O.I32, O.U8, O.I16, O.U32, O.I64, O.F32, O.F64])
@pytest.mark.parametrize("before", [True, False])
def test_inverse_scale_offset_division_bit_exact(ctx, dst, before):
    """(v - offset) / scale (write_helpers.rs:15-17) uses a reciprocal refined once per work item plus the division's own
    residual correction; it must equal the IEEE division of the oracle for every operand: random bit patterns (infinities,
    NaNs, denormals, huge and tiny magnitudes), quotients next to integers, and scales of every magnitude."""
    if not before and dst != O.F64:
        pytest.skip("a transform applied after the cast works on the target type: f64 only")
    rng = np.random.default_rng(77 + dst)
    n = 400_000
    ol, pl = util.layouts([("v", O.F64, 0)])
    olt, plt = util.layouts([("v", dst, 0)])
    scales = [0.001, 0.01, 3.0, -0.37, 1e-300, 1e300, 5e-324, 2.0 ** -1022, 1.0, float(np.nextafter(1.0, 2.0)), 1e-7, 123456.789,
              float("inf"), float("nan")]
    for k, s in enumerate(scales):
        o = [0.0, 500000.0, -1e-3][k % 3]
        vals = rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64).copy()
        q = rng.integers(-2 ** 31 - 5, 2 ** 31 + 5, n // 4).astype(np.float64)           # quotients at / next to integers
        with np.errstate(all="ignore"):
            near = q * s + o
            vals[: n // 4] = near
            vals[n // 4: n // 2] = np.nextafter(near, np.inf)
            vals[n // 2: 3 * n // 4] = np.nextafter(near, -np.inf)
        osrc = O.OBuffer(ol, n, True)
        osrc.columns[0][:] = vals.view(np.uint8)
        psrc = util.to_pb(osrc, pl)
        t = pb.InvScaleOffset(s, (o, o, o))
        ocv = O.OConverter(ol, olt, with_default=False)
        ocv.set_custom_mapping_with_transformation(("v", O.F64), ("v", dst), O.F64, util.oracle_transform(t), before)
        pcv = BufferLayoutConverter.for_layouts(pl, plt)
        pcv.set_custom_mapping_with_transformation(util.PointAttributeDefinition("v", O.F64, 0), util.PointAttributeDefinition("v", dst, 0), t, before)
        odst = ocv.convert(osrc, True)
        pdst = pcv.convert(psrc, HashMapBuffer)
        torch.cuda.synchronize()
        util.assert_buffers_match(odst, pdst, f"scale {s!r} offset {o}")


def test_schedule_autotune_is_result_neutral(ctx):
    """"convert.autotune": the first large conversion of a plan shape times alternative schedules on a prefix of its own
    range and keeps the fastest.  Schedules only move work between warps: the bytes (and the fused accumulators: AABB,
    out-of-range count) must be those of the untuned run and of the oracle; the trial runs must not leak into the accumulators."""
    n = 4_500_000
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[: 20 * n] = O.gen_las_fmt0_records(0, n)
    psrc = util.to_pb(osrc, pl)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    odst = O.OConverter.las_default(ol, olt, scale, offset).convert(osrc, True)
    pcv = pb.get_default_las_converter(pl, plt, scale, offset)
    ref = HashMapBuffer(plt, n, "cuda")
    b0 = pcv.convert_into_range_with_bounds(psrc, range(0, n), ref, range(0, n))
    ctx.set_param("convert.autotune", 1)
    try:
        for _ in range(2):  # first call tunes, second reuses the tuned schedule
            got = HashMapBuffer(plt, n, "cuda")
            b1 = pcv.convert_into_range_with_bounds(psrc, range(0, n), got, range(0, n))
            torch.cuda.synchronize()
            util.assert_buffers_match(odst, got, "autotuned C2")
            assert b1 == b0
        # write direction with the out-of-range counter and a packed 35 B source
        aos = BufferLayoutConverter.for_layouts_with_default(plt, plt).convert(got, VectorBuffer)
        oaos = O.OConverter(olt, olt, with_default=True).convert(odst, False)
        util.assert_buffers_match(oaos, aos, "autotuned columnar -> interleaved")
        t = pb.InvScaleOffset(scale, offset)
        owr = O.OConverter(olt, ol, with_default=True)
        owr.set_custom_mapping_with_transformation(("Position3D", O.VEC3F64), ("LASLocalPosition", O.VEC3I32), O.VEC3F64,
                                                   util.oracle_transform(t), True)
        oback = owr.convert(oaos, False)
        wr = BufferLayoutConverter.for_layouts_with_default(plt, pl)
        wr.set_custom_mapping_with_transformation(A.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, t, True)
        for _ in range(2):
            back = VectorBuffer(pl, n, "cuda")
            oor = wr.convert_into(aos, back, count_out_of_range=True)
            torch.cuda.synchronize()
            assert oor == 0
            util.assert_buffers_match(oback, back, "autotuned write direction")
    finally:
        ctx.set_param("convert.autotune", 0)
