"""ctypes loader for the CPU oracle (oracle/pasture_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pasture_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpasture_oracle.so")

MAX_ATTRS = 48
NAME_LEN = 64

# PointAttributeDataType codes (point_layout.rs:23-68 declaration order)
U8, I8, U16, I16, U32, I32, U64, I64, F32, F64 = range(10)
VEC3U8, VEC3U16, VEC3F32, VEC3I32, VEC3F64, VEC4U8, BYTEARRAY, CUSTOM = range(10, 18)

T_NONE, T_SCALE_OFFSET, T_INV_SCALE_OFFSET, T_ADD, T_SHIFT_MASK = range(5)

OK = 0
ERR_ATTR_NOT_FOUND, ERR_NO_CONVERSION, ERR_TRANSFORM_DTYPE, ERR_LAYOUT_MISMATCH = -1, -2, -3, -4
ERR_RANGE, ERR_DUPLICATE_ATTR, ERR_OVERLAP, ERR_INVALID, ERR_TOO_FEW_POINTS, ERR_UNSUPPORTED = -5, -6, -7, -8, -9, -10

NP_DTYPES = {U8: np.uint8, I8: np.int8, U16: np.uint16, I16: np.int16, U32: np.uint32, I32: np.int32,
             U64: np.uint64, I64: np.int64, F32: np.float32, F64: np.float64}
VEC3_COMPONENT = {VEC3U8: U8, VEC3U16: U16, VEC3F32: F32, VEC3I32: I32, VEC3F64: F64}


class Member(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("dtype", C.c_uint32), ("extra_size", C.c_uint64),
                ("extra_align", C.c_uint64), ("offset", C.c_uint64), ("size", C.c_uint64)]


class Layout(C.Structure):
    _fields_ = [("n", C.c_uint32), ("m", Member * MAX_ATTRS), ("size", C.c_uint64), ("align", C.c_uint64)]


class Transform(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("shift", C.c_uint32), ("mask", C.c_uint64),
                ("s", C.c_double * 3), ("o", C.c_double * 3)]


class Mapping(C.Structure):
    _fields_ = [("target_idx", C.c_int32), ("source_idx", C.c_int32), ("has_converter", C.c_int32),
                ("has_transform", C.c_int32), ("apply_to_source", C.c_int32), ("_pad", C.c_int32),
                ("t", Transform)]


class Converter(C.Structure):
    _fields_ = [("from_layout", Layout), ("to_layout", Layout), ("n_mappings", C.c_uint32),
                ("mappings", Mapping * MAX_ATTRS)]


class Buffer(C.Structure):
    _fields_ = [("layout", C.POINTER(Layout)), ("columnar", C.c_int32), ("len", C.c_uint64),
                ("aos", C.c_void_p), ("columns", C.POINTER(C.c_void_p))]


class ProjOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("p", C.c_double * 12)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "pasture_oracle.c")
    if not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        build()
    L = C.CDLL(_SO)
    u64, u32, i32, dbl, vp = C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_void_p
    LP, CP, BP = C.POINTER(Layout), C.POINTER(Converter), C.POINTER(Buffer)
    sig = {
        "po_dtype_size": (u64, [u32, u64]),
        "po_dtype_min_alignment": (u64, [u32, u64]),
        "po_layout_init": (None, [LP]),
        "po_layout_add_attribute": (i32, [LP, C.c_char_p, u32, u64, u64, u64]),
        "po_layout_from_members_and_alignment": (i32, [LP, C.POINTER(Member), u32, u64]),
        "po_layout_index_by_name": (i32, [LP, C.c_char_p]),
        "po_layout_index_of": (i32, [LP, C.c_char_p, u32]),
        "po_layout_equal": (i32, [LP, LP]),
        "po_has_conversion": (i32, [u32, u32]),
        "po_convert_value": (i32, [u32, u32, vp, vp]),
        "po_converter_for_layouts": (i32, [CP, LP, LP, i32]),
        "po_converter_set_custom_mapping": (i32, [CP, C.c_char_p, u32, C.c_char_p, u32]),
        "po_converter_set_custom_mapping_with_transformation":
            (i32, [CP, C.c_char_p, u32, C.c_char_p, u32, u32, C.POINTER(Transform), i32]),
        "po_convert_into_range": (i32, [CP, BP, u64, u64, BP, u64, u64]),
        "po_convert_into_range_mt": (i32, [CP, BP, u64, u64, BP, u64, u64, i32]),
        "po_las_raw_layout": (i32, [i32, LP]),
        "po_las_default_layout": (i32, [i32, LP]),
        "po_las_default_converter": (i32, [CP, LP, LP, C.POINTER(dbl), C.POINTER(dbl)]),
        "po_las_write_position": (i32, [C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(C.c_int32)]),
        "po_las_write_points": (i32, [BP, i32, C.POINTER(dbl), C.POINTER(dbl), vp, vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(u64)]),
        "po_calculate_bounds": (i32, [BP, C.POINTER(dbl), C.POINTER(dbl)]),
        "po_minmax_attribute": (i32, [BP, C.c_char_p, u32, u32, vp, vp]),
        "po_expand_bits_by_3": (u64, [u64]),
        "po_reverse_bits": (u64, [u64]),
        "po_create_markers": (u64, [dbl, dbl, dbl, vp, u64]),
        "po_find_leaf": (None, [vp, vp, u64, vp, u64, vp, u64, vp]),
        "po_find_leaf_bsearch": (None, [vp, vp, u64, vp, u64, vp, u64, vp]),
        "po_voxelgrid_filter": (i32, [BP, dbl, dbl, dbl, BP, vp, i32]),
        "po_knn_bruteforce": (None, [vp, u64, vp, u64, u32, vp, vp]),
        "po_knn_kdtree": (i32, [vp, u64, u64, u64, u32, vp, vp, vp, vp, i32]),
        "po_compute_centroid": (None, [vp, u64, vp]),
        "po_compute_covariance": (i32, [vp, u64, vp]),
        "po_solve_plane_parameter": (None, [vp, vp, vp]),
        "po_normal_estimation": (i32, [vp, u64, vp, vp]),
        "po_compute_normals": (i32, [vp, u64, u32, vp, vp]),
        "po_reproject": (None, [C.POINTER(ProjOp), u32, vp, vp, u64]),
        "po_pipeline_epsg4326_to_3309": (u32, [C.POINTER(ProjOp)]),
        "po_splitmix64": (u64, [u64, u64]),
        "po_ransac_model_from_samples": (None, [i32, vp, vp, vp]),
        "po_ransac_distance": (dbl, [i32, vp, vp]),
        "po_ransac_rank_models": (None, [i32, vp, u64, vp, u64, dbl, vp]),
        "po_ransac_inliers": (u64, [i32, vp, u64, vp, dbl, vp]),
        "po_ransac_draw_samples": (None, [i32, u64, u64, u64, vp]),
        "po_gen_las_fmt0_records": (None, [vp, u64, u64, u64]),
        "po_gen_c1_points": (None, [vp, u64, u64, u64, C.POINTER(dbl)]),
        "po_gen_terrain_positions": (None, [vp, u64, u64, u64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


class OracleError(RuntimeError):
    def __init__(self, code, what=""):
        super().__init__(f"oracle 'panic' code {code} {what}")
        self.code = code


def _check(rc, what=""):
    if rc < 0:
        raise OracleError(rc, what)
    return rc


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OLayout:
    """PointLayout restatement (point_layout.rs:648-997)."""

    def __init__(self):
        self.c = Layout()
        lib().po_layout_init(C.byref(self.c))

    @classmethod
    def from_attributes(cls, attrs, packed=0):
        """attrs: iterable of (name, dtype) or (name, dtype, extra_size). packed=0 -> FieldAlignment::Default."""
        l = cls()
        for a in attrs:
            l.add_attribute(*a, packed=packed)
        return l

    @classmethod
    def from_members_and_alignment(cls, members, alignment):
        """members: iterable of (name, dtype, offset)"""
        arr = (Member * len(members))()
        for i, (name, dtype, offset) in enumerate(members):
            arr[i].name = name.encode()
            arr[i].dtype = dtype
            arr[i].offset = offset
        l = cls()
        _check(lib().po_layout_from_members_and_alignment(C.byref(l.c), arr, len(members), alignment))
        return l

    @classmethod
    def las_raw(cls, fmt):
        l = cls()
        _check(lib().po_las_raw_layout(fmt, C.byref(l.c)))
        return l

    @classmethod
    def las_default(cls, fmt):
        l = cls()
        _check(lib().po_las_default_layout(fmt, C.byref(l.c)))
        return l

    def add_attribute(self, name, dtype, extra_size=0, extra_align=0, packed=0):
        _check(lib().po_layout_add_attribute(C.byref(self.c), name.encode(), dtype, extra_size, extra_align, packed),
               f"add_attribute({name})")

    @property
    def n(self):
        return self.c.n

    @property
    def size(self):
        return self.c.size

    @property
    def align(self):
        return self.c.align

    def members(self):
        return [(self.c.m[i].name.decode(), self.c.m[i].dtype, self.c.m[i].offset, self.c.m[i].size)
                for i in range(self.c.n)]

    def index_by_name(self, name):
        return lib().po_layout_index_by_name(C.byref(self.c), name.encode())

    def __eq__(self, other):
        return bool(lib().po_layout_equal(C.byref(self.c), C.byref(other.c)))


class OBuffer:
    """VectorBuffer (columnar=False) / HashMapBuffer (columnar=True) memory over numpy (zero-filled like resize())."""

    def __init__(self, layout, n, columnar):
        self.layout = layout
        self.columnar = bool(columnar)
        self.len = int(n)
        if columnar:
            self.columns = [np.zeros(max(1, n * sz), dtype=np.uint8) for (_, _, _, sz) in layout.members()]
            self.aos = None
        else:
            self.aos = np.zeros(max(1, n * layout.size), dtype=np.uint8)
            self.columns = None
        self._make_c()

    def _make_c(self):
        self.c = Buffer()
        self.c.layout = C.pointer(self.layout.c)
        self.c.columnar = 1 if self.columnar else 0
        self.c.len = self.len
        if self.columnar:
            self._colptrs = (C.c_void_p * len(self.columns))(*[c.ctypes.data for c in self.columns])
            self.c.columns = C.cast(self._colptrs, C.POINTER(C.c_void_p))
            self.c.aos = None
        else:
            self.c.aos = self.aos.ctypes.data
            self.c.columns = None

    def attribute_bytes(self, idx):
        """(len, size) uint8 copy of attribute idx"""
        name, dtype, off, sz = self.layout.members()[idx]
        if self.columnar:
            return self.columns[idx][: self.len * sz].reshape(self.len, sz).copy()
        rec = self.aos[: self.len * self.layout.size].reshape(self.len, self.layout.size)
        return rec[:, off:off + sz].copy()

    def attribute(self, name):
        """typed numpy view-copy of an attribute by name (scalars -> (n,), vec3 -> (n,3))"""
        idx = self.layout.index_by_name(name)
        _, dtype, _, sz = self.layout.members()[idx]
        raw = np.ascontiguousarray(self.attribute_bytes(idx))
        if dtype in NP_DTYPES:
            return raw.view(NP_DTYPES[dtype]).reshape(self.len)
        if dtype in VEC3_COMPONENT:
            return raw.view(NP_DTYPES[VEC3_COMPONENT[dtype]]).reshape(self.len, 3)
        return raw

    def set_attribute(self, name, values):
        idx = self.layout.index_by_name(name)
        _, dtype, off, sz = self.layout.members()[idx]
        comp = NP_DTYPES.get(dtype) or NP_DTYPES[VEC3_COMPONENT[dtype]]
        raw = np.ascontiguousarray(np.asarray(values, dtype=comp)).view(np.uint8).reshape(self.len, sz)
        if self.columnar:
            self.columns[idx][: self.len * sz] = raw.reshape(-1)
        else:
            rec = self.aos[: self.len * self.layout.size].reshape(self.len, self.layout.size)
            rec[:, off:off + sz] = raw

    def set_len(self, n):
        self.len = int(n)
        self.c.len = self.len


class OConverter:
    """BufferLayoutConverter restatement (buffer_conversion.rs:98-663)."""

    def __init__(self, from_layout, to_layout, with_default=False):
        self.c = Converter()
        self.from_layout, self.to_layout = from_layout, to_layout
        _check(lib().po_converter_for_layouts(C.byref(self.c), C.byref(from_layout.c), C.byref(to_layout.c),
                                              1 if with_default else 0), "for_layouts")

    @classmethod
    def las_default(cls, raw, target, scale, offset):
        self = cls.__new__(cls)
        self.c = Converter()
        self.from_layout, self.to_layout = raw, target
        s = (C.c_double * 3)(*scale)
        o = (C.c_double * 3)(*offset)
        _check(lib().po_las_default_converter(C.byref(self.c), C.byref(raw.c), C.byref(target.c), s, o),
               "las_default_converter")
        return self

    def set_custom_mapping(self, from_attr, to_attr):
        _check(lib().po_converter_set_custom_mapping(C.byref(self.c), from_attr[0].encode(), from_attr[1],
                                                     to_attr[0].encode(), to_attr[1]), "set_custom_mapping")

    def set_custom_mapping_with_transformation(self, from_attr, to_attr, transform_dtype, transform,
                                               apply_to_source):
        _check(lib().po_converter_set_custom_mapping_with_transformation(
            C.byref(self.c), from_attr[0].encode(), from_attr[1], to_attr[0].encode(), to_attr[1],
            transform_dtype, C.byref(transform), 1 if apply_to_source else 0), "set_custom_mapping_with_transformation")

    def convert_into_range(self, src, sb, se, dst, db, de, threads=0):
        if threads and threads > 1:
            _check(lib().po_convert_into_range_mt(C.byref(self.c), C.byref(src.c), sb, se, C.byref(dst.c), db, de,
                                                  threads), "convert_into_range_mt")
        else:
            _check(lib().po_convert_into_range(C.byref(self.c), C.byref(src.c), sb, se, C.byref(dst.c), db, de),
                   "convert_into_range")

    def convert_into(self, src, dst):
        self.convert_into_range(src, 0, src.len, dst, 0, src.len)

    def convert(self, src, columnar):
        dst = OBuffer(self.to_layout, src.len, columnar)
        self.convert_into(src, dst)
        return dst


def make_transform(kind, s=(1.0, 1.0, 1.0), o=(0.0, 0.0, 0.0), shift=0, mask=0):
    t = Transform()
    t.kind = kind
    t.shift = shift
    t.mask = mask
    for i in range(3):
        t.s[i] = s[i]
        t.o[i] = o[i]
    return t


def convert_value(from_dtype, to_dtype, value_bytes):
    src = np.frombuffer(bytes(value_bytes), dtype=np.uint8).copy()
    dst = np.zeros(32, dtype=np.uint8)
    _check(lib().po_convert_value(from_dtype, to_dtype, _ptr(src), _ptr(dst)), "convert_value")
    return dst[: lib().po_dtype_size(to_dtype, 0)].tobytes()


def las_write_points(src, fmt, scale, offset):
    """-> (records uint8 (n, record_size), counts16, bmin, bmax, panics)"""
    raw = OLayout.las_raw(fmt)
    out = np.zeros(max(1, src.len * raw.size), dtype=np.uint8)
    counts = np.zeros(16, dtype=np.uint64)
    mn, mx, panics = (C.c_double * 3)(), (C.c_double * 3)(), C.c_uint64(0)
    _check(lib().po_las_write_points(C.byref(src.c), fmt, (C.c_double * 3)(*scale), (C.c_double * 3)(*offset), _ptr(out),
                                     _ptr(counts), mn, mx, C.byref(panics)), "las_write_points")
    return out[: src.len * raw.size].reshape(src.len, raw.size), counts, np.array(mn[:]), np.array(mx[:]), int(panics.value)


def filter_into(src, predicate, dst):
    """HashMapBuffer::filter_into (pasture-core/src/containers/point_buffer.rs:1086-1136): for every attribute, the
    values of the points whose index satisfies `predicate` are written, in order, to the front of `dst`.
    -> number of matches. Raises like the reference panics (layout mismatch, target too small)."""
    if [m[:2] + m[3:] for m in src.layout.members()] != [m[:2] + m[3:] for m in dst.layout.members()] or \
            src.layout.size != dst.layout.size:
        raise OracleError(-4, "PointLayouts must match")  # :1092
    keep = [i for i in range(src.len) if predicate(i)]
    if dst.len < len(keep):
        raise OracleError(-5, "buffer.len() must be at least as large as the number of predicate matches")  # :1097
    for a, (_, _, off, sz) in enumerate(src.layout.members()):
        vals = src.attribute_bytes(a)[keep] if keep else np.zeros((0, sz), np.uint8)
        if dst.columnar:  # :1104-1119
            dst.columns[a][: len(keep) * sz] = vals.reshape(-1)
        else:  # :1120-1134
            rec = dst.aos[: dst.len * dst.layout.size].reshape(dst.len, dst.layout.size)
            rec[: len(keep), off:off + sz] = vals
    return len(keep)


def _set_attribute_bytes(buf, idx, rows, raw):
    _, _, off, sz = buf.layout.members()[idx]
    if buf.columnar:
        buf.columns[idx][: buf.len * sz].reshape(buf.len, sz)[rows] = raw
    else:
        buf.aos[: buf.len * buf.layout.size].reshape(buf.len, buf.layout.size)[rows, off:off + sz] = raw


def _convert_column(name, raw, from_dtype, to_dtype):
    """(n, size(from)) uint8 -> (n, size(to)) uint8 through get_converter_for_attributes (attribute_conversion.rs:160-271)"""
    n = raw.shape[0]
    if from_dtype == to_dtype or n == 0:
        return raw
    sl, tl = OLayout.from_attributes([(name, from_dtype)]), OLayout.from_attributes([(name, to_dtype)])
    src = OBuffer(sl, n, True)
    src.columns[0][: raw.size] = raw.reshape(-1)
    return OConverter(sl, tl).convert(src, True).attribute_bytes(0)


PNTS_SEMANTIC_DTYPES = {"Position3D": VEC3F32, "ColorRGB": VEC3U8, "ColorRGBA": VEC4U8, "Normal": VEC3F32}  # pnts_writer.rs:110-117


def pnts_read_into(body, file_attrs, first, count, dst, rtc_center=None):
    """PntsReader::read_into (pasture-io/src/tiles3d/pnts_reader.rs:294-367). body: uint8 array holding the file,
    file_attrs: [(name, dtype, byte offset of the array)]; points [first, first+count) -> dst[0, count).
    rtc_center: apply_rtc_center_offset (:247-283) over the WHOLE buffer."""
    body = np.frombuffer(bytes(body), dtype=np.uint8)
    for name, dtype, off in file_attrs:
        ti = dst.layout.index_by_name(name)
        if ti < 0:
            continue  # :308
        ssz = lib().po_dtype_size(dtype, 0)
        raw = body[off + first * ssz: off + (first + count) * ssz].reshape(count, ssz)
        _set_attribute_bytes(dst, ti, slice(0, count), _convert_column(name, raw, dtype, dst.layout.members()[ti][1]))
    pi = dst.layout.index_by_name("Position3D")
    if rtc_center is not None and pi >= 0:
        dtype = dst.layout.members()[pi][1]
        c = np.asarray(rtc_center, dtype=np.float64)
        p = dst.attribute("Position3D")
        if dtype == VEC3F32:
            p = (p.astype(np.float64) + c).astype(np.float32)  # :265-273
        elif dtype == VEC3F64:
            p = p + c  # :275-278
        else:
            raise OracleError(-10, "Unsupported datatype for POSITION_3D attribute")
        dst.set_attribute("Position3D", p)


def pnts_feature_table_body(src):
    """PntsWriter::write + write_feature_table_body (pnts_writer.rs:353-401, :310-341) for one buffer:
    -> ([(name, dtype, byte offset)], body bytes)"""
    attrs, parts, off = [], [], 0
    for i, (name, dtype, _, _) in enumerate(src.layout.members()):
        if name not in PNTS_SEMANTIC_DTYPES:
            continue
        to = PNTS_SEMANTIC_DTYPES[name]
        raw = _convert_column(name, src.attribute_bytes(i), dtype, to).reshape(-1)
        attrs.append((name, to, off))
        pad = (-raw.size) % 8
        parts += [raw, np.zeros(pad, np.uint8)]
        off += raw.size + pad
    return attrs, (np.concatenate(parts) if parts else np.zeros(0, np.uint8)).tobytes()


def calculate_bounds(buf):
    mn = (C.c_double * 3)()
    mx = (C.c_double * 3)()
    rc = _check(lib().po_calculate_bounds(C.byref(buf.c), mn, mx))
    return (np.array(mn[:]), np.array(mx[:])) if rc == 1 else None


def minmax_attribute(buf, name, attr_dtype, view_dtype=None):
    view_dtype = attr_dtype if view_dtype is None else view_dtype
    mn = np.zeros(32, dtype=np.uint8)
    mx = np.zeros(32, dtype=np.uint8)
    rc = _check(lib().po_minmax_attribute(C.byref(buf.c), name.encode(), attr_dtype, view_dtype, _ptr(mn), _ptr(mx)))
    if rc == 0:
        return None
    sz = lib().po_dtype_size(attr_dtype, 0)
    comp = NP_DTYPES.get(attr_dtype) or NP_DTYPES[VEC3_COMPONENT[attr_dtype]]
    return mn[:sz].view(comp).copy(), mx[:sz].view(comp).copy()


def create_markers(bmin, bmax, leaf):
    n = lib().po_create_markers(bmin, bmax, leaf, None, 0)
    out = np.zeros(max(1, n), dtype=np.float64)
    lib().po_create_markers(bmin, bmax, leaf, _ptr(out), n)
    return out[:n]


def find_leaf(p, mx, my, mz, bsearch=False):
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros(3, dtype=np.uint64)
    f = lib().po_find_leaf_bsearch if bsearch else lib().po_find_leaf
    f(_ptr(p), _ptr(mx), len(mx), _ptr(my), len(my), _ptr(mz), len(mz), _ptr(out))
    return tuple(int(x) for x in out)


def voxelgrid_filter(src, leaf, dst_layout, columnar=True, use_sort=False):
    dst = OBuffer(dst_layout, src.len, columnar)
    keys = np.zeros((max(1, src.len), 3), dtype=np.uint64)
    _check(lib().po_voxelgrid_filter(C.byref(src.c), leaf[0], leaf[1], leaf[2], C.byref(dst.c), _ptr(keys),
                                     1 if use_sort else 0), "voxelgrid_filter")
    dst.len = int(dst.c.len)
    return dst, keys[: dst.len].copy()


def bits_for(count):
    """bits needed for indices 0..count-1 (at least 1): the width of one axis in a packed voxel key"""
    b = 1
    while (1 << b) < count:
        b += 1
    return b


def voxel_partials(pts, gmin, gmax, leaf):
    """One shard's contribution to a sharded voxel grid, restated with the reference's own pieces: markers
    of the GLOBAL box (voxel_grid.rs:54-79), find_leaf per point (:22-51), position sums in point order (:339-379
    without the final division).  -> (packed keys int64 ascending, counts int32, sums [V,3], bits, cells)"""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    markers = [create_markers(gmin[c], gmax[c], leaf[c]) for c in range(3)]
    bits = [bits_for(len(m)) for m in markers]
    keys = np.zeros(len(pts), dtype=np.uint64)
    for i, p in enumerate(pts):
        ix, iy, iz = find_leaf(p, *markers)
        keys[i] = (ix << (bits[1] + bits[2])) | (iy << bits[2]) | iz
    k, c, s = merge_partials(keys.astype(np.int64), np.ones(len(pts), np.int32), pts)
    return k, c, s, bits, [len(m) for m in markers]


def merge_partials(keys, counts, sums):
    """equal keys are added in the order given (stable), sequential f64 adds"""
    keys = np.asarray(keys, dtype=np.int64)
    order = np.argsort(keys, kind="stable")
    uniq, starts, cnt = np.unique(keys[order], return_index=True, return_counts=True)
    out_c = np.zeros(len(uniq), dtype=np.int32)
    out_s = np.zeros((len(uniq), 3), dtype=np.float64)
    for v, (s0, n) in enumerate(zip(starts, cnt)):
        acc = np.zeros(3)
        tot = 0
        for j in order[s0:s0 + n]:
            acc = acc + sums[j]
            tot += int(counts[j])
        out_s[v] = acc
        out_c[v] = tot
    return uniq, out_c, out_s


def knn_bruteforce(pts, queries, k):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    idx = np.zeros((len(queries), k), dtype=np.uint32)
    d2 = np.zeros((len(queries), k), dtype=np.float64)
    lib().po_knn_bruteforce(_ptr(pts), len(pts), _ptr(queries), len(queries), k, _ptr(idx), _ptr(d2))
    return idx, d2


def knn_kdtree(pts, k, q_begin=0, q_end=None, with_distances=True, threads=None):
    """exact kNN of the cloud's own points [q_begin, q_end) over a kd-tree (all host threads by default)"""
    import os
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    q_end = len(pts) if q_end is None else q_end
    nq = max(0, q_end - q_begin)
    idx = np.zeros((nq, k), dtype=np.uint32)
    d2 = np.zeros((nq, k), dtype=np.float64) if with_distances else None
    _check(lib().po_knn_kdtree(_ptr(pts), len(pts), q_begin, q_end, k, _ptr(idx), _ptr(d2) if with_distances else None, None, None,
                               threads or os.cpu_count() or 1), "knn_kdtree")
    return (idx, d2) if with_distances else idx


def compute_normals_kdtree(pts, k, q_begin=0, q_end=None, threads=None):
    """compute_normals (normal_estimation.rs:79-130) with the kd-tree kNN: (normals, curvature) of the points [q_begin, q_end)"""
    import os
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    q_end = len(pts) if q_end is None else q_end
    nq = max(0, q_end - q_begin)
    normals = np.zeros((nq, 3))
    curv = np.zeros(nq)
    _check(lib().po_knn_kdtree(_ptr(pts), len(pts), q_begin, q_end, k, None, None, _ptr(normals), _ptr(curv),
                               threads or os.cpu_count() or 1), "compute_normals_kdtree")
    return normals, curv


def compute_centroid(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros(3)
    lib().po_compute_centroid(_ptr(pts), len(pts), _ptr(out))
    return out


def compute_covariance(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros(9)
    _check(lib().po_compute_covariance(_ptr(pts), len(pts), _ptr(out)), "compute_covariance_matrix")
    return out.reshape(3, 3)


def solve_plane_parameter(cov):
    cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(9)
    n = np.zeros(3)
    c = np.zeros(1)
    lib().po_solve_plane_parameter(_ptr(cov), _ptr(n), _ptr(c))
    return n, float(c[0])


def compute_normals(pts, k):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    normals = np.zeros((len(pts), 3))
    curv = np.zeros(len(pts))
    _check(lib().po_compute_normals(_ptr(pts), len(pts), k, _ptr(normals), _ptr(curv)), "compute_normals")
    return normals, curv


def pipeline_epsg4326_to_3309():
    ops = (ProjOp * 8)()
    n = lib().po_pipeline_epsg4326_to_3309(ops)
    return ops, n


PROJ_AFFINE, PROJ_GEODETIC_TO_ECEF, PROJ_ECEF_TO_GEODETIC, PROJ_ALBERS_FWD, PROJ_SET_Z, PROJ_WEBMERC_FWD, PROJ_TMERC_FWD = range(1, 8)
PROJ_DEG2RAD_LATLON, PROJ_RAD2DEG_LATLON, PROJ_TMERC_INV, PROJ_WEBMERC_INV = 8, 9, 10, 11
WGS84, GRS80 = (6378137.0, 298.257223563), (6378137.0, 298.257222101)


def make_pipeline(steps):
    """[(kind, [params...]), ...] -> (ops array, n)"""
    ops = (ProjOp * 8)()
    for i, (kind, params) in enumerate(steps):
        ops[i].kind = kind
        for j, v in enumerate(params):
            ops[i].p[j] = float(v)
    return ops, len(steps)


def tmerc_step(ellipsoid, lat0_deg, lon0_deg, k0, fe, fn, inverse=False):
    import math
    return (PROJ_TMERC_INV if inverse else PROJ_TMERC_FWD,
            [ellipsoid[0], ellipsoid[1], math.radians(lat0_deg), math.radians(lon0_deg), k0, fe, fn])


def utm_step(zone, south=False, ellipsoid=WGS84, inverse=False):
    """UTM zone `zone` (EPSG:326zz / 327zz on WGS 84, 258zz on GRS 1980): central meridian 6 zone - 183 deg, k0 0.9996,
    FE 500 000 m, FN 0 / 10 000 000 m (south)"""
    return tmerc_step(ellipsoid, 0.0, 6.0 * zone - 183.0, 0.9996, 500000.0, 10000000.0 if south else 0.0, inverse)


def helmert_step(tx, ty, tz, rx_as, ry_as, rz_as, ds_ppm):
    """7-parameter Position Vector transformation (EPSG method 1033, GN7-2 4.3.3) as an affine step"""
    import math
    s = math.pi / (180.0 * 3600.0)
    rx, ry, rz, m = rx_as * s, ry_as * s, rz_as * s, 1.0 + ds_ppm * 1e-6
    return (PROJ_AFFINE, [m, -m * rz, m * ry, m * rz, m, -m * rx, -m * ry, m * rx, m, tx, ty, tz])


def reproject(ops, n_ops, xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.zeros_like(xyz)
    lib().po_reproject(ops, n_ops, _ptr(xyz), _ptr(out), len(xyz))
    return out


def ransac_draw_samples(kind, n, n_models, seed):
    s = np.zeros((n_models, 3 if kind == 0 else 2), dtype=np.uint64)
    lib().po_ransac_draw_samples(kind, n, n_models, seed, _ptr(s))
    return s


def ransac_models(kind, pts, samples):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    samples = np.ascontiguousarray(samples, dtype=np.uint64)
    w = 4 if kind == 0 else 6
    m = np.zeros((len(samples), w))
    for h in range(len(samples)):
        row = np.zeros(w)
        lib().po_ransac_model_from_samples(kind, _ptr(pts), _ptr(samples[h].copy()), _ptr(row))
        m[h] = row
    return m


def ransac_rank_models(kind, pts, models, threshold):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    models = np.ascontiguousarray(models, dtype=np.float64)
    r = np.zeros(len(models), dtype=np.uint64)
    lib().po_ransac_rank_models(kind, _ptr(pts), len(pts), _ptr(models), len(models), C.c_double(threshold), _ptr(r))
    return r


def ransac_inliers(kind, pts, model, threshold):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    model = np.ascontiguousarray(model, dtype=np.float64)
    idx = np.zeros(max(1, len(pts)), dtype=np.uint64)
    k = lib().po_ransac_inliers(kind, _ptr(pts), len(pts), _ptr(model), C.c_double(threshold), _ptr(idx))
    return idx[:k].copy()


def ransac(kind, pts, threshold, n_models, seed):
    """ransac_{plane,line}_serial (segmentation.rs:240-255, :350-368) with the seeded draw: (model, ranking, indices)"""
    samples = ransac_draw_samples(kind, len(pts), n_models, seed)
    models = ransac_models(kind, pts, samples)
    ranks = ransac_rank_models(kind, pts, models, threshold)
    best = len(ranks) - 1 - int(np.argmax(ranks[::-1]))  # Iterator::max_by returns the last maximum
    return models[best], int(ranks[best]), ransac_inliers(kind, pts, models[best], threshold)


def gen_las_fmt0_records(first, n, seed=42):
    out = np.zeros(max(1, 20 * n), dtype=np.uint8)
    lib().po_gen_las_fmt0_records(_ptr(out), first, n, seed)
    return out[: 20 * n]


def gen_c1_points(first, n, seed=42, offset=(0.0, 0.0, 0.0)):
    out = np.zeros(max(1, 35 * n), dtype=np.uint8)
    o = (C.c_double * 3)(*offset)
    lib().po_gen_c1_points(_ptr(out), first, n, seed, o)
    return out[: 35 * n]


def gen_terrain_positions(first, n, seed=42):
    out = np.zeros((max(1, n), 3), dtype=np.float64)
    lib().po_gen_terrain_positions(_ptr(out), first, n, seed)
    return out[:n]
