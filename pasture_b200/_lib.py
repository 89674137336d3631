"""ctypes binding of libpasture_b200.so (the C ABI declared in include/pasture_b200.h).

The library is the product; this module only loads it.  If it is missing the import fails loudly:
there is no Python/CPU fallback for any compute entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpasture_b200.so")

MAX_NAME = 64


class Attr(C.Structure):
    _fields_ = [("name", C.c_char * MAX_NAME), ("dtype", C.c_uint32), ("_pad", C.c_uint32),
                ("extra_size", C.c_uint64), ("extra_align", C.c_uint64), ("offset", C.c_uint64),
                ("size", C.c_uint64)]


class BufferDesc(C.Structure):
    _fields_ = [("layout", C.c_void_p), ("kind", C.c_int32), ("memspace", C.c_int32), ("len", C.c_uint64),
                ("aos", C.c_void_p), ("columns", C.POINTER(C.c_void_p))]


class Transform(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("shift", C.c_uint32), ("mask", C.c_uint64), ("s", C.c_double * 3),
                ("o", C.c_double * 3)]


class LasHeader(C.Structure):
    _fields_ = [("version_major", C.c_uint8), ("version_minor", C.c_uint8), ("point_format", C.c_uint8),
                ("is_compressed", C.c_uint8), ("record_length", C.c_uint16), ("header_size", C.c_uint16),
                ("offset_to_point_data", C.c_uint32), ("number_of_vlrs", C.c_uint32), ("extra_bytes", C.c_uint32),
                ("_pad", C.c_uint32), ("number_of_points", C.c_uint64), ("scale", C.c_double * 3),
                ("offset", C.c_double * 3), ("min", C.c_double * 3), ("max", C.c_double * 3)]


class LasWriteStats(C.Structure):
    _fields_ = [("out_of_range", C.c_uint64), ("points_by_return", C.c_uint64 * 16), ("has_bounds", C.c_int32),
                ("_pad", C.c_int32), ("bounds_min", C.c_double * 3), ("bounds_max", C.c_double * 3)]


class VoxelPartialsDesc(C.Structure):
    _fields_ = [("len", C.c_uint64), ("keys", C.c_void_p), ("counts", C.c_void_p), ("sums", C.c_void_p),
                ("bits_x", C.c_uint32), ("bits_y", C.c_uint32), ("bits_z", C.c_uint32), ("_pad", C.c_uint32),
                ("cells", C.c_uint64 * 3)]


class VoxelAttrPartialsDesc(C.Structure):
    _fields_ = [("n_columns", C.c_uint32), ("n_modes", C.c_uint32), ("columns", C.c_void_p), ("column_is_max", C.c_uint8 * 64),
                ("mode_len", C.c_uint64 * 48), ("mode_keys", C.c_void_p * 48), ("mode_counts", C.c_void_p * 48)]


class ProjOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("p", C.c_double * 12)]


vp, u64, u32, i32, i64, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int64, C.c_double
PVP = C.POINTER(C.c_void_p)
PD = C.POINTER(C.c_double)
BD = C.POINTER(BufferDesc)

# name -> (restype, argtypes); must list every symbol declared in include/pasture_b200.h
SIGNATURES = {
    "pb200_abi_version": (i32, []),
    "pb200_last_error": (C.c_char_p, []),
    "pb200_kernel_launch_count": (u64, []),
    "pb200_ctx_create": (i32, [i32, PVP]),
    "pb200_ctx_set_stream": (i32, [vp, vp]),
    "pb200_ctx_get_stream": (vp, [vp]),
    "pb200_ctx_synchronize": (i32, [vp]),
    "pb200_ctx_trim": (i32, [vp]),
    "pb200_ctx_set_param": (i32, [vp, C.c_char_p, i64]),
    "pb200_ctx_profile_read": (i32, [vp, C.c_char_p, u64]),
    "pb200_ctx_destroy": (None, [vp]),
    "pb200_ctx_bind_host_thread": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "pb200_host_alloc": (i32, [u64, PVP]),
    "pb200_host_free": (i32, [vp]),
    "pb200_device_alloc": (i32, [vp, u64, PVP]),
    "pb200_device_free": (i32, [vp, vp]),
    "pb200_memcpy_h2d": (i32, [vp, vp, vp, u64]),
    "pb200_memcpy_d2h": (i32, [vp, vp, vp, u64]),
    "pb200_memcpy_d2d": (i32, [vp, vp, vp, u64]),
    "pb200_memset_device": (i32, [vp, vp, i32, u64]),
    "pb200_dtype_size": (u64, [u32, u64]),
    "pb200_dtype_min_alignment": (u64, [u32, u64]),
    "pb200_layout_create": (i32, [PVP]),
    "pb200_layout_add_attribute": (i32, [vp, C.c_char_p, u32, u64, u64, u64]),
    "pb200_layout_from_members_and_alignment": (i32, [C.POINTER(Attr), u32, u64, PVP]),
    "pb200_layout_clone": (i32, [vp, PVP]),
    "pb200_layout_num_attributes": (u32, [vp]),
    "pb200_layout_get_attribute": (i32, [vp, u32, C.POINTER(Attr)]),
    "pb200_layout_size_of_point_entry": (u64, [vp]),
    "pb200_layout_alignment": (u64, [vp]),
    "pb200_layout_index_by_name": (i32, [vp, C.c_char_p]),
    "pb200_layout_index_of": (i32, [vp, C.c_char_p, u32]),
    "pb200_layout_equal": (i32, [vp, vp]),
    "pb200_layout_compare_without_offsets": (i32, [vp, vp]),
    "pb200_layout_destroy": (None, [vp]),
    "pb200_las_raw_layout": (i32, [i32, PVP]),
    "pb200_las_default_layout": (i32, [i32, PVP]),
    "pb200_converter_create": (i32, [vp, vp, vp, i32, PVP]),
    "pb200_converter_set_custom_mapping": (i32, [vp, C.c_char_p, u32, C.c_char_p, u32]),
    "pb200_converter_set_custom_mapping_with_transformation":
        (i32, [vp, C.c_char_p, u32, C.c_char_p, u32, u32, C.POINTER(Transform), i32]),
    "pb200_las_default_converter": (i32, [vp, vp, vp, PD, PD, PVP]),
    "pb200_converter_set_packed_mapping": (i32, [vp, C.c_char_p, u32, u32, C.POINTER(C.c_char_p), C.POINTER(u32), C.POINTER(u32)]),
    "pb200_las_parse_header": (i32, [vp, u64, C.POINTER(LasHeader)]),
    "pb200_las_read_points": (i32, [vp, vp, u64, u64, u64, BD, u64]),
    "pb200_las_write_points": (i32, [vp, BD, u64, u64, i32, PD, PD, vp, C.c_int32, C.POINTER(LasWriteStats)]),
    "pb200_converter_num_mappings": (u32, [vp]),
    "pb200_converter_convert_into_range": (i32, [vp, BD, u64, u64, BD, u64, u64, C.POINTER(u64)]),
    "pb200_converter_convert_fresh_range": (i32, [vp, BD, u64, u64, BD, u64, u64, C.POINTER(u64)]),
    "pb200_converter_describe_schedule": (i32, [vp, BD, u64, u64, BD, u64, i32, C.c_char_p, u64]),
    "pb200_converter_convert_into": (i32, [vp, BD, BD, C.POINTER(u64)]),
    "pb200_converter_convert_into_range_with_bounds": (i32, [vp, BD, u64, u64, BD, u64, u64, PD, PD, C.POINTER(i32)]),
    "pb200_converter_convert_into_range_with_bounds_device": (i32, [vp, BD, u64, u64, BD, u64, u64, vp]),
    "pb200_converter_destroy": (None, [vp]),
    "pb200_transform_attribute": (i32, [vp, BD, C.c_char_p, u32, C.POINTER(Transform)]),
    "pb200_view_attribute_with_conversion": (i32, [vp, BD, C.c_char_p, u32, vp]),
    "pb200_pnts_compatible_layout": (i32, [vp, u64, vp, C.POINTER(u32), C.POINTER(u64)]),
    "pb200_pnts_read_points": (i32, [vp, vp, u64, vp, u32, u64, u64, BD, vp]),
    "pb200_pnts_write_points": (i32, [vp, BD, vp, u64]),
    "pb200_ransac_rank_samples": (i32, [vp, BD, i32, vp, u64, C.c_double, vp, vp]),
    "pb200_ransac_rank_models": (i32, [vp, BD, i32, vp, u64, C.c_double, vp]),
    "pb200_ransac_inliers": (i32, [vp, BD, i32, vp, C.c_double, vp, u64, C.POINTER(u64)]),
    "pb200_ransac": (i32, [vp, BD, i32, C.c_double, u64, u64, vp, C.POINTER(u64), vp, u64]),
    "pb200_filter_into": (i32, [vp, BD, vp, BD, C.POINTER(u64)]),
    "pb200_calculate_bounds": (i32, [vp, BD, PD, PD, C.POINTER(i32)]),
    "pb200_minmax_attribute": (i32, [vp, BD, C.c_char_p, u32, vp, vp, C.POINTER(i32)]),
    "pb200_minmax_attribute_partial": (i32, [vp, BD, C.c_char_p, u32, vp, vp, C.POINTER(i32)]),
    "pb200_expand_bits_by_3": (u64, [u64]),
    "pb200_morton_codes": (i32, [vp, BD, PD, PD, vp]),
    "pb200_voxelgrid_filter": (i32, [vp, BD, dbl, dbl, dbl, vp, C.c_int32, C.c_int32, PVP]),
    "pb200_result_buffer_desc": (i32, [vp, BD]),
    "pb200_result_buffer_voxel_keys": (i32, [vp, vp]),
    "pb200_result_buffer_destroy": (None, [vp]),
    "pb200_radix_sort_u64": (i32, [vp, vp, vp, u64, i32, i32]),
    "pb200_comm_create": (i32, [vp, i32, i32, PVP]),
    "pb200_comm_handle": (i32, [vp, vp]),
    "pb200_comm_connect": (i32, [vp, vp]),
    "pb200_comm_exchange_ptr": (i32, [vp, PVP]),
    "pb200_comm_connect_ptrs": (i32, [vp, PVP]),
    "pb200_comm_check": (i32, [vp]),
    "pb200_comm_destroy": (None, [vp]),
    "pb200_converter_convert_into_range_with_global_bounds": (i32, [vp, BD, u64, u64, BD, u64, u64, vp, vp]),
    "pb200_voxelgrid_partials": (i32, [vp, BD, dbl, dbl, dbl, PD, PD, PVP]),
    "pb200_voxelgrid_merge_partials": (i32, [vp, vp, vp, vp, u64, u32, u32, u32, PVP]),
    "pb200_voxel_partials_get": (i32, [vp, C.POINTER(VoxelPartialsDesc)]),
    "pb200_voxelgrid_partials_layout": (i32, [vp, BD, dbl, dbl, dbl, PD, PD, vp, PVP]),
    "pb200_voxel_partials_get_attrs": (i32, [vp, C.POINTER(VoxelAttrPartialsDesc)]),
    "pb200_voxelgrid_merge_partials_layout": (i32, [vp, vp, C.POINTER(VoxelPartialsDesc), C.POINTER(VoxelAttrPartialsDesc), PVP]),
    "pb200_voxel_partials_centroids": (i32, [vp, vp]),
    "pb200_voxel_partials_destroy": (None, [vp]),
    "pb200_knn": (i32, [vp, BD, u32, vp, vp]),
    "pb200_knn_range": (i32, [vp, BD, u32, u64, u64, vp, vp]),
    "pb200_compute_normals_range": (i32, [vp, BD, u32, u64, u64, vp, vp]),
    "pb200_radius_search": (i32, [vp, BD, dbl, u32, vp, vp]),
    "pb200_compute_normals": (i32, [vp, BD, u32, vp, vp]),
    "pb200_proj_pipeline_for_crs": (i32, [C.c_char_p, C.c_char_p, C.POINTER(ProjOp), u32]),
    "pb200_proj_op_tmerc": (i32, [dbl, dbl, dbl, dbl, dbl, dbl, dbl, i32, C.POINTER(ProjOp)]),
    "pb200_proj_op_helmert": (i32, [dbl, dbl, dbl, dbl, dbl, dbl, dbl, i32, C.POINTER(ProjOp)]),
    "pb200_reproject": (i32, [vp, BD, BD, C.POINTER(ProjOp), u32]),
    "pb200_synth_las_fmt0_records": (i32, [vp, vp, u64, u64, u64]),
    "pb200_synth_terrain_positions": (i32, [vp, vp, u64, u64, u64]),
}

_lib = None


class PastureB200Error(RuntimeError):
    """A contract violation (a `panic!` in the reference) or a CUDA failure reported by the C ABI."""

    def __init__(self, code, message):
        super().__init__(f"[pb200 error {code}] {message}")
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C pasture_b200/csrc). pasture_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)  # AttributeError here = header/library mismatch
        f.restype = res
        f.argtypes = args
    if L.pb200_abi_version() != 1:
        raise ImportError("libpasture_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise PastureB200Error(rc, lib().pb200_last_error().decode(errors="replace"))
    return rc
