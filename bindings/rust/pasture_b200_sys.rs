//! Raw FFI declarations for libpasture_b200 (include/pasture_b200.h, ABI version 1).
//!
//! UNVERIFIED: there is no Rust toolchain in the image this repository is built in, so this file has never been
//! compiled. It mirrors the C header declaration by declaration; `tests/test_abi_host.py` verifies the header against
//! the shared library, not this file.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const PB200_ABI_VERSION: c_int = 1;
pub const PB200_MAX_NAME: usize = 64;

#[repr(C)] pub struct pb200_ctx { _private: [u8; 0] }
#[repr(C)] pub struct pb200_layout { _private: [u8; 0] }
#[repr(C)] pub struct pb200_converter { _private: [u8; 0] }
#[repr(C)] pub struct pb200_result_buffer { _private: [u8; 0] }

/// PointAttributeDataType codes (pasture-core/src/layout/point_layout.rs:23-68, declaration order)
pub mod dtype {
    pub const U8: u32 = 0; pub const I8: u32 = 1; pub const U16: u32 = 2; pub const I16: u32 = 3;
    pub const U32: u32 = 4; pub const I32: u32 = 5; pub const U64: u32 = 6; pub const I64: u32 = 7;
    pub const F32: u32 = 8; pub const F64: u32 = 9; pub const VEC3U8: u32 = 10; pub const VEC3U16: u32 = 11;
    pub const VEC3F32: u32 = 12; pub const VEC3I32: u32 = 13; pub const VEC3F64: u32 = 14; pub const VEC4U8: u32 = 15;
    pub const BYTEARRAY: u32 = 16; pub const CUSTOM: u32 = 17;
}

pub const PB200_OK: c_int = 0;
pub const PB200_ERR_ATTR_NOT_FOUND: c_int = -1;
pub const PB200_ERR_NO_CONVERSION: c_int = -2;
pub const PB200_ERR_TRANSFORM_DTYPE: c_int = -3;
pub const PB200_ERR_LAYOUT_MISMATCH: c_int = -4;
pub const PB200_ERR_RANGE: c_int = -5;
pub const PB200_ERR_DUPLICATE_ATTR: c_int = -6;
pub const PB200_ERR_OVERLAP: c_int = -7;
pub const PB200_ERR_INVALID: c_int = -8;
pub const PB200_ERR_TOO_FEW_POINTS: c_int = -9;
pub const PB200_ERR_UNSUPPORTED: c_int = -10;
pub const PB200_ERR_NO_DEVICE: c_int = -100;
pub const PB200_ERR_CUDA: c_int = -101;
pub const PB200_ERR_OOM: c_int = -102;

pub const PB200_INTERLEAVED: i32 = 0;
pub const PB200_COLUMNAR: i32 = 1;
pub const PB200_HOST: i32 = 0;
pub const PB200_DEVICE: i32 = 1;

#[repr(C)] #[derive(Clone, Copy)]
pub struct pb200_attr {
    pub name: [c_char; PB200_MAX_NAME], pub dtype: u32, pub _pad: u32,
    pub extra_size: u64, pub extra_align: u64, pub offset: u64, pub size: u64,
}

#[repr(C)]
pub struct pb200_buffer_desc {
    pub layout: *const pb200_layout, pub kind: i32, pub memspace: i32, pub len: u64,
    pub aos: *mut c_void, pub columns: *mut *mut c_void,
}

pub const PB200_T_NONE: u32 = 0; pub const PB200_T_SCALE_OFFSET: u32 = 1; pub const PB200_T_INV_SCALE_OFFSET: u32 = 2;
pub const PB200_T_ADD: u32 = 3; pub const PB200_T_SHIFT_MASK: u32 = 4;

#[repr(C)] #[derive(Clone, Copy)]
pub struct pb200_transform { pub kind: u32, pub shift: u32, pub mask: u64, pub s: [f64; 3], pub o: [f64; 3] }

#[repr(C)] #[derive(Clone, Copy)]
pub struct pb200_las_header {
    pub version_major: u8, pub version_minor: u8, pub point_format: u8, pub is_compressed: u8,
    pub record_length: u16, pub header_size: u16,
    pub offset_to_point_data: u32, pub number_of_vlrs: u32, pub extra_bytes: u32, pub _pad: u32,
    pub number_of_points: u64,
    pub scale: [f64; 3], pub offset: [f64; 3], pub min: [f64; 3], pub max: [f64; 3],
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct pb200_las_write_stats {
    pub out_of_range: u64, pub points_by_return: [u64; 16], pub has_bounds: i32, pub _pad: i32,
    pub bounds_min: [f64; 3], pub bounds_max: [f64; 3],
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct pb200_proj_op { pub kind: u32, pub _pad: u32, pub p: [f64; 12] }

#[link(name = "pasture_b200")]
extern "C" {
    pub fn pb200_abi_version() -> c_int;
    pub fn pb200_last_error() -> *const c_char;
    pub fn pb200_kernel_launch_count() -> u64;
    pub fn pb200_ctx_create(device: c_int, out: *mut *mut pb200_ctx) -> c_int;
    pub fn pb200_ctx_set_stream(ctx: *mut pb200_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn pb200_ctx_get_stream(ctx: *mut pb200_ctx) -> *mut c_void;
    pub fn pb200_ctx_synchronize(ctx: *mut pb200_ctx) -> c_int;
    pub fn pb200_ctx_trim(ctx: *mut pb200_ctx) -> c_int;
    pub fn pb200_ctx_set_param(ctx: *mut pb200_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn pb200_ctx_destroy(ctx: *mut pb200_ctx);
    pub fn pb200_ctx_profile_read(ctx: *mut pb200_ctx, out: *mut c_char, capacity: u64) -> c_int;
    pub fn pb200_ctx_bind_host_thread(ctx: *mut pb200_ctx, numa_node_out: *mut c_int, n_cpus_out: *mut c_int) -> c_int;
    pub fn pb200_host_alloc(bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn pb200_host_free(p: *mut c_void) -> c_int;
    pub fn pb200_device_alloc(ctx: *mut pb200_ctx, bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn pb200_device_free(ctx: *mut pb200_ctx, p: *mut c_void) -> c_int;
    pub fn pb200_memcpy_h2d(ctx: *mut pb200_ctx, dst: *mut c_void, src: *const c_void, bytes: u64) -> c_int;
    pub fn pb200_memcpy_d2h(ctx: *mut pb200_ctx, dst: *mut c_void, src: *const c_void, bytes: u64) -> c_int;
    pub fn pb200_memcpy_d2d(ctx: *mut pb200_ctx, dst: *mut c_void, src: *const c_void, bytes: u64) -> c_int;
    pub fn pb200_memset_device(ctx: *mut pb200_ctx, dst: *mut c_void, value: c_int, bytes: u64) -> c_int;

    pub fn pb200_dtype_size(dtype: u32, extra_size: u64) -> u64;
    pub fn pb200_dtype_min_alignment(dtype: u32, extra_align: u64) -> u64;
    pub fn pb200_layout_create(out: *mut *mut pb200_layout) -> c_int;
    pub fn pb200_layout_add_attribute(l: *mut pb200_layout, name: *const c_char, dtype: u32, extra_size: u64,
                                      extra_align: u64, packed_n: u64) -> c_int;
    pub fn pb200_layout_from_members_and_alignment(members: *const pb200_attr, n: u32, type_alignment: u64,
                                                   out: *mut *mut pb200_layout) -> c_int;
    pub fn pb200_layout_clone(l: *const pb200_layout, out: *mut *mut pb200_layout) -> c_int;
    pub fn pb200_layout_num_attributes(l: *const pb200_layout) -> u32;
    pub fn pb200_layout_get_attribute(l: *const pb200_layout, index: u32, out: *mut pb200_attr) -> c_int;
    pub fn pb200_layout_size_of_point_entry(l: *const pb200_layout) -> u64;
    pub fn pb200_layout_alignment(l: *const pb200_layout) -> u64;
    pub fn pb200_layout_index_by_name(l: *const pb200_layout, name: *const c_char) -> c_int;
    pub fn pb200_layout_index_of(l: *const pb200_layout, name: *const c_char, dtype: u32) -> c_int;
    pub fn pb200_layout_equal(a: *const pb200_layout, b: *const pb200_layout) -> c_int;
    pub fn pb200_layout_compare_without_offsets(a: *const pb200_layout, b: *const pb200_layout) -> c_int;
    pub fn pb200_layout_destroy(l: *mut pb200_layout);
    pub fn pb200_las_raw_layout(point_format: c_int, out: *mut *mut pb200_layout) -> c_int;
    pub fn pb200_las_default_layout(point_format: c_int, out: *mut *mut pb200_layout) -> c_int;

    pub fn pb200_converter_create(ctx: *mut pb200_ctx, from: *const pb200_layout, to: *const pb200_layout,
                                  with_default: c_int, out: *mut *mut pb200_converter) -> c_int;
    pub fn pb200_converter_set_custom_mapping(cv: *mut pb200_converter, from_name: *const c_char, from_dtype: u32,
                                              to_name: *const c_char, to_dtype: u32) -> c_int;
    pub fn pb200_converter_set_custom_mapping_with_transformation(cv: *mut pb200_converter, from_name: *const c_char,
        from_dtype: u32, to_name: *const c_char, to_dtype: u32, transform_dtype: u32, t: *const pb200_transform,
        apply_to_source: c_int) -> c_int;
    pub fn pb200_las_default_converter(ctx: *mut pb200_ctx, raw: *const pb200_layout, target: *const pb200_layout,
                                       scale: *const f64, offset: *const f64, out: *mut *mut pb200_converter) -> c_int;
    pub fn pb200_converter_set_packed_mapping(cv: *mut pb200_converter, to_name: *const c_char, to_dtype: u32, n_sources: u32,
        from_names: *const *const c_char, masks: *const u32, shifts: *const u32) -> c_int;
    pub fn pb200_las_parse_header(file_bytes: *const c_void, size: u64, out: *mut pb200_las_header) -> c_int;
    pub fn pb200_las_read_points(ctx: *mut pb200_ctx, file_bytes: *const c_void, size: u64, first_point: u64, count: u64,
                                 dst: *const pb200_buffer_desc, dst_begin: u64) -> c_int;
    pub fn pb200_las_write_points(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, begin: u64, end: u64, point_format: c_int,
        scale: *const f64, offset: *const f64, out_records: *mut c_void, out_memspace: i32, stats: *mut pb200_las_write_stats) -> c_int;
    pub fn pb200_converter_num_mappings(cv: *const pb200_converter) -> u32;
    pub fn pb200_converter_convert_into_range(cv: *mut pb200_converter, src: *const pb200_buffer_desc, src_begin: u64,
        src_end: u64, dst: *const pb200_buffer_desc, dst_begin: u64, dst_end: u64, out_of_range_count: *mut u64) -> c_int;
    pub fn pb200_converter_convert_fresh_range(cv: *mut pb200_converter, src: *const pb200_buffer_desc, src_begin: u64, src_end: u64,
                                               dst: *const pb200_buffer_desc, dst_begin: u64, dst_end: u64,
                                               out_of_range_count: *mut u64) -> c_int;
    /// Host-side description of the tile schedule (no device needed; `cv` may come from `pb200_converter_create(null, ..)`).
    pub fn pb200_converter_describe_schedule(cv: *const pb200_converter, src: *const pb200_buffer_desc, src_begin: u64, src_end: u64,
                                             dst: *const pb200_buffer_desc, dst_begin: u64, fresh_target: i32,
                                             out: *mut core::ffi::c_char, capacity: u64) -> i32;
    pub fn pb200_converter_convert_into(cv: *mut pb200_converter, src: *const pb200_buffer_desc,
                                        dst: *const pb200_buffer_desc, out_of_range_count: *mut u64) -> c_int;
    pub fn pb200_converter_convert_into_range_with_bounds(cv: *mut pb200_converter, src: *const pb200_buffer_desc,
        src_begin: u64, src_end: u64, dst: *const pb200_buffer_desc, dst_begin: u64, dst_end: u64,
        out_min: *mut f64, out_max: *mut f64, is_some: *mut c_int) -> c_int;
    pub fn pb200_converter_convert_into_range_with_bounds_device(cv: *mut pb200_converter, src: *const pb200_buffer_desc,
        src_begin: u64, src_end: u64, dst: *const pb200_buffer_desc, dst_begin: u64, dst_end: u64,
        device_minmax6: *mut f64) -> c_int;
    pub fn pb200_converter_destroy(cv: *mut pb200_converter);
    pub fn pb200_transform_attribute(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, name: *const c_char, dtype: u32,
                                     t: *const pb200_transform) -> c_int;
    pub fn pb200_view_attribute_with_conversion(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, name: *const c_char,
                                                view_dtype: u32, out: *mut c_void) -> c_int;

    pub fn pb200_pnts_compatible_layout(point_layout: *const pb200_layout, num_points: u64, out_attrs: *mut pb200_attr, n_out: *mut u32,
                                        body_bytes: *mut u64) -> c_int;
    pub fn pb200_pnts_read_points(ctx: *mut pb200_ctx, body: *const c_void, body_size: u64, attrs: *const pb200_attr, n_attrs: u32,
                                  first_point: u64, count: u64, dst: *const pb200_buffer_desc, rtc_center: *const f64) -> c_int;
    pub fn pb200_pnts_write_points(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, body_out: *mut c_void, body_capacity: u64) -> c_int;
    pub fn pb200_ransac_rank_samples(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, kind: c_int, samples: *const u64, n_models: u64,
                                     distance_threshold: f64, models_out: *mut f64, rankings_out: *mut u64) -> c_int;
    pub fn pb200_ransac_rank_models(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, kind: c_int, models: *const f64, n_models: u64,
                                    distance_threshold: f64, rankings_out: *mut u64) -> c_int;
    pub fn pb200_ransac_inliers(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, kind: c_int, model: *const f64, distance_threshold: f64,
                                indices_out: *mut u64, capacity: u64, num_inliers: *mut u64) -> c_int;
    pub fn pb200_ransac(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, kind: c_int, distance_threshold: f64, num_of_iterations: u64,
                        seed: u64, model_out: *mut f64, ranking_out: *mut u64, indices_out: *mut u64, capacity: u64) -> c_int;
    pub fn pb200_filter_into(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, mask: *const u8, dst: *const pb200_buffer_desc,
                             num_matches: *mut u64) -> c_int;
    pub fn pb200_calculate_bounds(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, out_min: *mut f64,
                                  out_max: *mut f64, is_some: *mut c_int) -> c_int;
    pub fn pb200_minmax_attribute(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, name: *const c_char, dtype: u32,
                                  out_min: *mut c_void, out_max: *mut c_void, is_some: *mut c_int) -> c_int;
    pub fn pb200_minmax_attribute_partial(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, name: *const c_char, dtype: u32,
                                          out_min: *mut c_void, out_max: *mut c_void, is_some: *mut c_int) -> c_int;
    pub fn pb200_expand_bits_by_3(v: u64) -> u64;
    pub fn pb200_morton_codes(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, bmin: *const f64, bmax: *const f64,
                              codes_out: *mut u64) -> c_int;
    pub fn pb200_voxelgrid_filter(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, leaf_x: f64, leaf_y: f64,
        leaf_z: f64, dst_layout: *const pb200_layout, dst_kind: i32, dst_memspace: i32,
        out: *mut *mut pb200_result_buffer) -> c_int;
    pub fn pb200_result_buffer_desc(r: *const pb200_result_buffer, out: *mut pb200_buffer_desc) -> c_int;
    pub fn pb200_result_buffer_voxel_keys(r: *const pb200_result_buffer, keys_out: *mut u64) -> c_int;
    pub fn pb200_result_buffer_destroy(r: *mut pb200_result_buffer);
    pub fn pb200_knn(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, k: u32, idx_out: *mut u32, d2_out: *mut f64) -> c_int;
    pub fn pb200_knn_range(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, k: u32, first_query: u64, n_queries: u64,
                           idx_out: *mut u32, d2_out: *mut f64) -> c_int;
    pub fn pb200_compute_normals_range(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, k: u32, first_query: u64, n_queries: u64,
                                       normals_out: *mut f64, curvature_out: *mut f64) -> c_int;
    pub fn pb200_radius_search(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, radius: f64, max_neighbors: u32,
                               idx_out: *mut u32, counts_out: *mut u32) -> c_int;
    pub fn pb200_compute_normals(ctx: *mut pb200_ctx, buf: *const pb200_buffer_desc, k: u32, normals_out: *mut f64,
                                 curvature_out: *mut f64) -> c_int;
    pub fn pb200_proj_pipeline_for_crs(source_crs: *const c_char, target_crs: *const c_char, ops: *mut pb200_proj_op,
                                       cap: u32) -> c_int;
    pub fn pb200_proj_op_tmerc(a: f64, inv_f: f64, lat0_deg: f64, lon0_deg: f64, k0: f64, false_easting: f64, false_northing: f64,
                               inverse: c_int, out: *mut pb200_proj_op) -> c_int;
    pub fn pb200_proj_op_helmert(tx: f64, ty: f64, tz: f64, rx_arcsec: f64, ry_arcsec: f64, rz_arcsec: f64, ds_ppm: f64,
                                 coordinate_frame: c_int, out: *mut pb200_proj_op) -> c_int;
    pub fn pb200_reproject(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, dst_or_null: *const pb200_buffer_desc,
                           ops: *const pb200_proj_op, n_ops: u32) -> c_int;
    pub fn pb200_synth_las_fmt0_records(ctx: *mut pb200_ctx, device_out: *mut c_void, first_index: u64, n: u64, seed: u64) -> c_int;
    pub fn pb200_synth_terrain_positions(ctx: *mut pb200_ctx, device_out: *mut c_void, first_index: u64, n: u64, seed: u64) -> c_int;
    pub fn pb200_radix_sort_u64(ctx: *mut pb200_ctx, keys: *mut u64, vals: *mut u32, n: u64, begin_bit: c_int, end_bit: c_int) -> c_int;
    // ---- peer-memory communicator + sharded voxel grid ----
    pub fn pb200_comm_create(ctx: *mut pb200_ctx, rank: c_int, world: c_int, out: *mut *mut pb200_comm) -> c_int;
    pub fn pb200_comm_handle(c: *mut pb200_comm, handle_out: *mut c_void) -> c_int;
    pub fn pb200_comm_connect(c: *mut pb200_comm, handles: *const c_void) -> c_int;
    pub fn pb200_comm_exchange_ptr(c: *mut pb200_comm, device_ptr_out: *mut *mut c_void) -> c_int;
    pub fn pb200_comm_connect_ptrs(c: *mut pb200_comm, peer_ptrs: *const *mut c_void) -> c_int;
    pub fn pb200_comm_check(c: *mut pb200_comm) -> c_int;
    pub fn pb200_comm_destroy(c: *mut pb200_comm);
    pub fn pb200_converter_convert_into_range_with_global_bounds(cv: *mut pb200_converter, src: *const pb200_buffer_desc,
        src_begin: u64, src_end: u64, dst: *const pb200_buffer_desc, dst_begin: u64, dst_end: u64, comm: *mut pb200_comm,
        device_minmax6: *mut f64) -> c_int;
    pub fn pb200_voxelgrid_partials(ctx: *mut pb200_ctx, src: *const pb200_buffer_desc, leaf_x: f64, leaf_y: f64, leaf_z: f64,
        global_min: *const f64, global_max: *const f64, out: *mut *mut pb200_voxel_partials) -> c_int;
    pub fn pb200_voxelgrid_merge_partials(ctx: *mut pb200_ctx, keys: *const u64, counts: *const u32, sums: *const f64, m: u64,
        bits_x: u32, bits_y: u32, bits_z: u32, out: *mut *mut pb200_voxel_partials) -> c_int;
    pub fn pb200_voxel_partials_get(p: *const pb200_voxel_partials, out: *mut pb200_voxel_partials_desc) -> c_int;
    pub fn pb200_voxel_partials_centroids(p: *const pb200_voxel_partials, positions_out: *mut f64) -> c_int;
    pub fn pb200_voxel_partials_destroy(p: *mut pb200_voxel_partials);
}

#[repr(C)] pub struct pb200_comm { _private: [u8; 0] }
#[repr(C)] pub struct pb200_voxel_partials { _private: [u8; 0] }
#[repr(C)]
pub struct pb200_voxel_partials_desc {
    pub len: u64, pub keys: *const u64, pub counts: *const u32, pub sums: *const f64,
    pub bits_x: u32, pub bits_y: u32, pub bits_z: u32, pub _pad: u32, pub cells: [u64; 3],
}
