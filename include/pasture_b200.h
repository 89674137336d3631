/*
 * pasture_b200.h -- C ABI of the B200-native implementation of pasture's per-point hot path.
 *
 * Every entry point names the reference interface it stands in for (paths relative to the
 * igd-geo/pasture checkout, v0.5.0 @ 1b0b39c).  The reference has no FFI for this path (it is a
 * Rust trait surface); INTEGRATION.md shows the Rust-side binding a maintainer would add.
 *
 * Conventions
 *   - plain C types only; all handles are opaque; caller owns all point memory.
 *   - return value: 0 = ok, negative = error.  Contract violations that `panic!` in the reference
 *     come back as the PB200_ERR_* code documented on the function; pb200_last_error() gives the
 *     thread-local message.
 *   - handles are not thread-safe (the reference's BufferLayoutConverter is !Send as well,
 *     buffer_conversion.rs:14).  One CUDA stream per context.
 *   - buffers are described, not owned: a pb200_buffer_desc is the (base, offset, stride, size)
 *     addressing of containers/raw_attribute_view.rs:10-71 for a whole buffer.
 *   - memspace DEVICE: pointers are device pointers valid on the context's device; calls are
 *     asynchronous on the context's stream unless they return a value to the host.
 *     memspace HOST: the library stages chunks through device memory (H2D, kernel, D2H overlapped).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PB200_ERR_NO_DEVICE.
 */
#ifndef PASTURE_B200_H
#define PASTURE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 1
#define PB200_MAX_ATTRIBUTES 48
#define PB200_MAX_NAME 64

/* PointAttributeDataType, declaration order of pasture-core/src/layout/point_layout.rs:23-68 */
enum pb200_dtype {
    PB200_U8 = 0, PB200_I8 = 1, PB200_U16 = 2, PB200_I16 = 3, PB200_U32 = 4, PB200_I32 = 5,
    PB200_U64 = 6, PB200_I64 = 7, PB200_F32 = 8, PB200_F64 = 9,
    PB200_VEC3U8 = 10, PB200_VEC3U16 = 11, PB200_VEC3F32 = 12, PB200_VEC3I32 = 13, PB200_VEC3F64 = 14,
    PB200_VEC4U8 = 15, PB200_BYTEARRAY = 16, PB200_CUSTOM = 17
};

enum pb200_status {
    PB200_OK = 0,
    PB200_ERR_ATTR_NOT_FOUND = -1,  /* "Attribute not found" panics: buffer_conversion.rs:114,164,168 */
    PB200_ERR_NO_CONVERSION = -2,   /* "No conversion from .. to .." buffer_conversion.rs:383-388 */
    PB200_ERR_TRANSFORM_DTYPE = -3, /* assert_eq!(T::data_type(), ..) buffer_conversion.rs:209-213 */
    PB200_ERR_LAYOUT_MISMATCH = -4, /* buffer_conversion.rs:302-303 */
    PB200_ERR_RANGE = -5,           /* buffer_conversion.rs:304-306 ; reprojection.rs:212-214 */
    PB200_ERR_DUPLICATE_ATTR = -6,  /* point_layout.rs:783-788 */
    PB200_ERR_OVERLAP = -7,         /* point_layout.rs:737-743 */
    PB200_ERR_INVALID = -8,         /* bad argument / would panic (e.g. AABB min > max, math/bounds.rs:21-26) */
    PB200_ERR_TOO_FEW_POINTS = -9,  /* normal_estimation.rs:86-91,296-298 */
    PB200_ERR_UNSUPPORTED = -10,    /* voxel_grid.rs:452-459,682-687 ; raw_readers.rs:56 ; closure transforms */
    PB200_ERR_NO_DEVICE = -100,     /* no CUDA device / extension unusable: there is no CPU fallback */
    PB200_ERR_CUDA = -101,          /* CUDA runtime error (message in pb200_last_error) */
    PB200_ERR_OOM = -102
};

enum pb200_buffer_kind { PB200_INTERLEAVED = 0, PB200_COLUMNAR = 1 }; /* VectorBuffer / HashMapBuffer */
enum pb200_memspace { PB200_HOST = 0, PB200_DEVICE = 1 };

typedef struct pb200_ctx pb200_ctx;
typedef struct pb200_layout pb200_layout;
typedef struct pb200_converter pb200_converter;
typedef struct pb200_result_buffer pb200_result_buffer;

/* PointAttributeMember (point_layout.rs:353-357) */
typedef struct pb200_attr {
    char name[PB200_MAX_NAME];
    uint32_t dtype;
    uint32_t _pad;
    uint64_t extra_size;  /* ByteArray(n) length / Custom size */
    uint64_t extra_align; /* Custom min_alignment */
    uint64_t offset;
    uint64_t size;
} pb200_attr;

/* A borrowed point buffer: BorrowedBuffer + InterleavedBuffer|ColumnarBuffer (point_buffer.rs:17-654).
 * INTERLEAVED: `aos` holds len * size_of_point_entry bytes (VectorBuffer / ExternalMemoryBuffer).
 * COLUMNAR  : `columns[i]` holds len * size(attribute i) bytes, layout order (HashMapBuffer). */
typedef struct pb200_buffer_desc {
    const pb200_layout* layout;
    int32_t kind;     /* pb200_buffer_kind */
    int32_t memspace; /* pb200_memspace */
    uint64_t len;
    void* aos;
    void** columns;
} pb200_buffer_desc;

/* The closures that exist in-tree, enumerated (closures cannot cross an FFI):
 *   SCALE_OFFSET      Vec3f64/f64: (v*s)+o          Vec3f32/f32: ((v as f64*s)+o) as f32   pasture-io/src/las/raw_readers.rs:42-55
 *   INV_SCALE_OFFSET  Vec3f64/f64: (v-o)/s                                                 pasture-io/src/las/write_helpers.rs:15-17
 *   ADD               Vec3f64/f64: v+o              Vec3f32/f32: (v as f64+o) as f32       pasture-io/src/tiles3d/pnts_reader.rs:265-277
 *   SHIFT_MASK        u8/u16/u32/u64: (v >> shift) & mask                                  pasture-io/src/las/raw_readers.rs:61-164
 * No FMA contraction anywhere: results are bit-identical to the Rust closures. */
enum pb200_transform_kind {
    PB200_T_NONE = 0, PB200_T_SCALE_OFFSET = 1, PB200_T_INV_SCALE_OFFSET = 2, PB200_T_ADD = 3, PB200_T_SHIFT_MASK = 4
};
typedef struct pb200_transform {
    uint32_t kind;
    uint32_t shift;
    uint64_t mask;
    double s[3];
    double o[3];
} pb200_transform;

/* ---- library / context ------------------------------------------------------------------------ */
int pb200_abi_version(void);
const char* pb200_last_error(void);
/* number of kernels this library has launched in this process (bench.py `gpu_launches`) */
uint64_t pb200_kernel_launch_count(void);

/* A context must outlive every converter, communicator and result object created from it (they keep a pointer to it). */
int pb200_ctx_create(int device, pb200_ctx** out);
/* borrow an external CUDA stream (cudaStream_t as void*, e.g. torch.cuda.current_stream().cuda_stream) */
int pb200_ctx_set_stream(pb200_ctx* ctx, void* cuda_stream);
void* pb200_ctx_get_stream(pb200_ctx* ctx);
int pb200_ctx_synchronize(pb200_ctx* ctx);
/* hand the temporaries cached between calls (sort buffers, trees; stream-ordered pool) back to the driver */
int pb200_ctx_trim(pb200_ctx* ctx);
/* tuning knobs of the tile pipeline: "convert.tile_points", "convert.threads", "convert.stages",
 * "convert.ctas_per_sm", "convert.force_direct", "profile.phases" */
int pb200_ctx_set_param(pb200_ctx* ctx, const char* key, int64_t value);
/* Phase timer: after pb200_ctx_set_param(ctx, "profile.phases", 1) the library brackets the phases of its calls
 * (voxel.keys, sort.pass, voxel.reduce, knn.query, ...) with CUDA events on the context's stream.  This call synchronises,
 * writes one "name<TAB>milliseconds" line per recorded phase (call order) into out, clears the records and returns
 * their number.  Measurement tooling (bench.py other_configs); off by default. */
int pb200_ctx_profile_read(pb200_ctx* ctx, char* out, uint64_t capacity);
void pb200_ctx_destroy(pb200_ctx* ctx);
/* Bind the CALLING THREAD to the CPUs that are NUMA-local to the context's device (sysfs local_cpulist of the GPU's PCI
 * function, intersected with the thread's current affinity) so that the pinned host buffers it allocates afterwards, and
 * the copies it issues, stay on the GPU's socket.  *numa_node_out = the device's node (-1 unknown), *n_cpus_out = CPUs
 * bound (0: affinity left unchanged, e.g. no sysfs).  Multi-GPU hosts: call it once per rank before allocating HOST
 * buffers (what the reference's chunk loop, pasture-io/src/las/raw_readers.rs:309-348, reads from and writes to). */
int pb200_ctx_bind_host_thread(pb200_ctx* ctx, int* numa_node_out, int* n_cpus_out);

/* pinned host / device memory helpers (ExternalMemoryBuffer-style foreign memory, point_buffer.rs:1479-1503) */
int pb200_host_alloc(uint64_t bytes, void** out);
int pb200_host_free(void* p);
int pb200_device_alloc(pb200_ctx* ctx, uint64_t bytes, void** out);
int pb200_device_free(pb200_ctx* ctx, void* p);
int pb200_memcpy_h2d(pb200_ctx* ctx, void* dst_device, const void* src_host, uint64_t bytes);
int pb200_memcpy_d2h(pb200_ctx* ctx, void* dst_host, const void* src_device, uint64_t bytes);
int pb200_memcpy_d2d(pb200_ctx* ctx, void* dst_device, const void* src_device, uint64_t bytes);
int pb200_memset_device(pb200_ctx* ctx, void* dst_device, int value, uint64_t bytes);

/* ---- PointLayout (pasture-core/src/layout/point_layout.rs:648-997) ------------------------------ */
uint64_t pb200_dtype_size(uint32_t dtype, uint64_t extra_size);            /* :72-97   */
uint64_t pb200_dtype_min_alignment(uint32_t dtype, uint64_t extra_align);  /* :100-126 */
int pb200_layout_create(pb200_layout** out);                               /* PointLayout::default() :1011-1024 */
/* add_attribute :778-822. packed_n == 0 -> FieldAlignment::Default, else FieldAlignment::Packed(packed_n).
 * PB200_ERR_DUPLICATE_ATTR if the name is present. */
int pb200_layout_add_attribute(pb200_layout* l, const char* name, uint32_t dtype, uint64_t extra_size,
                               uint64_t extra_align, uint64_t packed_n);
/* from_members_and_alignment :719-759 (uses name, dtype, extra_*, offset of each attr) */
int pb200_layout_from_members_and_alignment(const pb200_attr* members, uint32_t n, uint64_t type_alignment,
                                            pb200_layout** out);
int pb200_layout_clone(const pb200_layout* l, pb200_layout** out);
uint32_t pb200_layout_num_attributes(const pb200_layout* l);
int pb200_layout_get_attribute(const pb200_layout* l, uint32_t index, pb200_attr* out);      /* at() :898 */
uint64_t pb200_layout_size_of_point_entry(const pb200_layout* l);                            /* :930 */
uint64_t pb200_layout_alignment(const pb200_layout* l);
int pb200_layout_index_by_name(const pb200_layout* l, const char* name);                     /* :887, -1 if absent */
int pb200_layout_index_of(const pb200_layout* l, const char* name, uint32_t dtype);          /* :951 */
int pb200_layout_equal(const pb200_layout* a, const pb200_layout* b);                        /* derived PartialEq */
int pb200_layout_compare_without_offsets(const pb200_layout* a, const pb200_layout* b);      /* :960-971 */
void pb200_layout_destroy(pb200_layout* l);
/* LAS layouts: pasture-io/src/las/las_layout.rs:64-125 (exact binary) and las_types.rs LasPointFormatN */
int pb200_las_raw_layout(int point_format, pb200_layout** out);
int pb200_las_default_layout(int point_format, pb200_layout** out);

/* ---- BufferLayoutConverter (pasture-core/src/layout/conversion/buffer_conversion.rs:98-663) ---- */
/* for_layouts :112 (with_default=0, PB200_ERR_ATTR_NOT_FOUND if a target attribute is missing in
 * `from`) / for_layouts_with_default :126 (with_default=1). PB200_ERR_NO_CONVERSION if a same-named
 * pair has no cast (attribute_conversion.rs:184-271). */
int pb200_converter_create(pb200_ctx* ctx, const pb200_layout* from, const pb200_layout* to, int with_default,
                           pb200_converter** out);
/* ctx may be NULL: a planning-only converter for pb200_converter_describe_schedule; every conversion with it returns
 * PB200_ERR_NO_DEVICE. */
/* Host-side description of the tile schedule a conversion of src[sb, se) into dst[db, ...) would run (no device needed;
 * the descriptors' pointers are only used as addresses).  Text, one line per work item:
 *   tiles tile_points=.. threads=.. stages=.. ctas_per_sm=.. smem=.. ops=.. items=.. load_first=..
 *   item warp=.. kind=copy|scalar|pack|zero src_type=.. dst_type=.. xf=.. bytes=.. group=.. src_rel=.. dst_rel=.. p0=.. p1=..
 * or "direct ops=.." when the plan does not fit the tile pipeline.  Returns the text length or a negative error code. */
int pb200_converter_describe_schedule(const pb200_converter* cv, const pb200_buffer_desc* src, uint64_t sb, uint64_t se,
                                      const pb200_buffer_desc* dst, uint64_t db, int fresh_target, char* out, uint64_t capacity);

/* set_custom_mapping :156-183 */
int pb200_converter_set_custom_mapping(pb200_converter* cv, const char* from_name, uint32_t from_dtype,
                                       const char* to_name, uint32_t to_dtype);
/* set_custom_mapping_with_transformation :194-234. `transform_dtype` plays the role of T: it must equal the
 * source dtype (apply_to_source != 0) or the target dtype, else PB200_ERR_TRANSFORM_DTYPE. */
int pb200_converter_set_custom_mapping_with_transformation(pb200_converter* cv, const char* from_name,
                                                           uint32_t from_dtype, const char* to_name,
                                                           uint32_t to_dtype, uint32_t transform_dtype,
                                                           const pb200_transform* t, int apply_to_source);
/* Bit-field packing (extension; the reference packs LAS flags per point in its writer, write_helpers.rs:26-51):
 * target = OR_k ((source_k & mask_k) << shift_k); sources are U8 attributes of the source layout, target is U8/U16. */
int pb200_converter_set_packed_mapping(pb200_converter* cv, const char* to_name, uint32_t to_dtype, uint32_t n_sources,
                                       const char* const* from_names, const uint32_t* masks, const uint32_t* shifts);
/* get_default_las_converter, pasture-io/src/las/raw_readers.rs:31-167 */
int pb200_las_default_converter(pb200_ctx* ctx, const pb200_layout* raw_las_layout, const pb200_layout* target,
                                const double scale[3], const double offset[3], pb200_converter** out);
uint32_t pb200_converter_num_mappings(const pb200_converter* cv);
/* convert_into_range :292-359. out_of_range_count (nullable): number of values for which a mapping with an
 * INV_SCALE_OFFSET transform followed by a float->int cast left the target integer range, i.e. the points
 * for which write_position_as_las_position (write_helpers.rs:15-17) would panic. Reading it synchronises. */
int pb200_converter_convert_into_range(pb200_converter* cv, const pb200_buffer_desc* src, uint64_t src_begin,
                                       uint64_t src_end, const pb200_buffer_desc* dst, uint64_t dst_begin,
                                       uint64_t dst_end, uint64_t* out_of_range_count);
/* BufferLayoutConverter::convert :242-259 on caller-owned memory: the reference creates the target with
 * `OutBuffer::new_from_layout` + `resize` (zero fill, point_buffer.rs:833-837) and converts into it, so the previous content
 * of the target range is irrelevant.  This entry point has those semantics for [dst_begin, dst_end): EVERY byte of the
 * range is written -- mapped attributes with their converted values, everything else (unmapped attributes, padding)
 * with zero -- and nothing is read back, which saves the read-modify-write of partially mapped interleaved records that
 * convert_into_range needs (it must preserve unmapped bytes).  The target may be uninitialised memory. */
int pb200_converter_convert_fresh_range(pb200_converter* cv, const pb200_buffer_desc* src, uint64_t src_begin,
                                        uint64_t src_end, const pb200_buffer_desc* dst, uint64_t dst_begin,
                                        uint64_t dst_end, uint64_t* out_of_range_count);
/* convert_into :268-283 */
int pb200_converter_convert_into(pb200_converter* cv, const pb200_buffer_desc* src, const pb200_buffer_desc* dst,
                                 uint64_t* out_of_range_count);
/* convert_into_range fused with calculate_bounds (pasture-algorithms/src/bounds.rs:30-54) over the produced
 * target POSITION_3D (Vec3f64) values: saves the separate 24 B/point pass. is_some=0 for an
 * empty range or a target without a mapped Vec3f64 "Position3D". Synchronises. */
int pb200_converter_convert_into_range_with_bounds(pb200_converter* cv, const pb200_buffer_desc* src,
                                                   uint64_t src_begin, uint64_t src_end,
                                                   const pb200_buffer_desc* dst, uint64_t dst_begin,
                                                   uint64_t dst_end, double out_min[3], double out_max[3],
                                                   int* is_some);
/* same, but leaves [minx,miny,minz,-maxx,-maxy,-maxz] in a device array (6 doubles) without synchronising,
 * ready for one ncclAllReduce(min) across point-range shards. Empty range: +MAX / +MAX.
 * Returns 1 if bounds were tracked (the target has a mapped Vec3f64 "Position3D"), 0 if not, < 0 on error. */
int pb200_converter_convert_into_range_with_bounds_device(pb200_converter* cv, const pb200_buffer_desc* src,
                                                          uint64_t src_begin, uint64_t src_end,
                                                          const pb200_buffer_desc* dst, uint64_t dst_begin,
                                                          uint64_t dst_end, double* device_minmax6);
void pb200_converter_destroy(pb200_converter* cv);

/* BorrowedMutBuffer::transform_attribute (point_buffer.rs:391-404) with an enumerated transform, in place */
int pb200_transform_attribute(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype,
                              const pb200_transform* t);
/* AttributeViewConverting (buffer_views.rs:533-650): materialise attribute `name` as `view_dtype`
 * into out (len * size(view_dtype) bytes, same memspace as buf) */
int pb200_view_attribute_with_conversion(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name,
                                         uint32_t view_dtype, void* out);

/* HashMapBuffer::filter_into (pasture-core/src/containers/point_buffer.rs:1086-1136) with the index predicate given as
 * a byte mask (len entries, non-zero = keep, same memory space as `src`). Keeps the order of the points. Writes the
 * number of kept points to *num_matches; PB200_ERR_RANGE if dst->len is smaller, PB200_ERR_LAYOUT_MISMATCH if the
 * layouts differ. Works for interleaved and columnar buffers on both sides. */
int pb200_filter_into(pb200_ctx* ctx, const pb200_buffer_desc* src, const uint8_t* mask, const pb200_buffer_desc* dst,
                      uint64_t* num_matches);

/* ---- reductions ----------------------------------------------------------------------------------- */
/* calculate_bounds, pasture-algorithms/src/bounds.rs:11-85. is_some=0 <=> None. All-NaN input ->
 * PB200_ERR_INVALID (AABB::from_min_max panics). */
int pb200_calculate_bounds(pb200_ctx* ctx, const pb200_buffer_desc* buf, double out_min[3], double out_max[3],
                           int* is_some);
/* minmax_attribute<T>, pasture-algorithms/src/minmax.rs:13-51; out_min/out_max hold one value of `dtype`.
 * (name, dtype) must be in the layout (the reference's converting branch always panics, buffer_views.rs:549). */
int pb200_minmax_attribute(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype,
                           void* out_min, void* out_max, int* is_some);
/* The fold of minmax_attribute over a shard that is NOT the beginning of the cloud ("minmax ints: same with the
 * attribute's dtype"): the reference seeds (min, max) with the cloud's FIRST value and then only replaces on strict
 * comparisons (math/minmax.rs:62-96), so later NaNs are ignored while a NaN seed sticks.  This entry point is the
 * continuation: NaN never enters, and a float component without any non-NaN value comes back as (+MAX, -MAX).  The
 * ranks' partial results combine by min / max; the rank holding point 0 applies the seed rule (pb200_minmax_attribute on
 * its shard, or the sharding helper's explicit seed exchange). */
int pb200_minmax_attribute_partial(pb200_ctx* ctx, const pb200_buffer_desc* buf, const char* name, uint32_t dtype,
                                   void* out_min, void* out_max, int* is_some);
/* expand_bits_by_3, pasture-core/src/math/bitmanip.rs:2-10 (host) and 63-bit Morton codes of positions
 * quantised to 21 bits/axis inside [bmin,bmax] (device; codes_out: len u64 in buf's memspace) */
uint64_t pb200_expand_bits_by_3(uint64_t v);
int pb200_morton_codes(pb200_ctx* ctx, const pb200_buffer_desc* buf, const double bmin[3], const double bmax[3],
                       uint64_t* codes_out);

/* ---- voxel grid, pasture-algorithms/src/voxel_grid.rs:109-165 ------------------------------------- */
/* Result: a library-owned buffer with `dst_layout`, one point per occupied voxel in (ix,iy,iz)
 * lexicographic order. Mode ties (nondeterministic in the reference, :320-328) resolve to the smallest
 * value. PB200_ERR_UNSUPPORTED for waveform / non-builtin target attributes (:452-459,:682-687). */
int pb200_voxelgrid_filter(pb200_ctx* ctx, const pb200_buffer_desc* src, double leaf_x, double leaf_y,
                           double leaf_z, const pb200_layout* dst_layout, int32_t dst_kind, int32_t dst_memspace,
                           pb200_result_buffer** out);
int pb200_result_buffer_desc(const pb200_result_buffer* r, pb200_buffer_desc* out);
/* per-voxel (ix,iy,iz) keys as 3 x u64, host memory, output order */
int pb200_result_buffer_voxel_keys(const pb200_result_buffer* r, uint64_t* keys_out);
void pb200_result_buffer_destroy(pb200_result_buffer* r);

/* ---- radix sort (K8) -----------------------------------------------------------------------------------
 * Stable LSD radix sort of n device-resident u64 keys by their bits [begin_bit, end_bit), optionally carrying a u32
 * payload (vals may be NULL); in place from the caller's point of view.  The sort behind the voxel grid and the LBVH. */
int pb200_radix_sort_u64(pb200_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t n, int begin_bit, int end_bit);

/* ---- peer-memory communicator --------------------------------------------------------
 * The global AABB of a cloud sharded over the GPUs of one box WITHOUT a separate collective: the last CTA of the fused
 * convert kernel stores its six min/max keys straight into every peer's exchange buffer over NVLink, signals, waits
 * for the peers' stores and reduces.  Each rank creates a communicator, publishes its 64-byte CUDA-IPC handle
 * (one process per GPU: exchange them with any host-side all-gather) or its raw device pointer (one process driving
 * several GPUs), connects, and then calls the _with_global_bounds conversion in lockstep with its peers. */
#define PB200_COMM_HANDLE_BYTES 64
typedef struct pb200_comm pb200_comm;
int pb200_comm_create(pb200_ctx* ctx, int rank, int world, pb200_comm** out);
int pb200_comm_handle(pb200_comm* c, void* handle_out /* PB200_COMM_HANDLE_BYTES */);
int pb200_comm_connect(pb200_comm* c, const void* handles /* world x PB200_COMM_HANDLE_BYTES, rank order */);
int pb200_comm_exchange_ptr(pb200_comm* c, void** device_ptr_out);
int pb200_comm_connect_ptrs(pb200_comm* c, void* const* peer_ptrs /* world device pointers, rank order */);
/* synchronises the context's stream; PB200_ERR_CUDA if a peer failed to arrive within the kernel's 10 s timeout */
int pb200_comm_check(pb200_comm* c);
void pb200_comm_destroy(pb200_comm* c);
/* convert_into_range + fused AABB of the produced POSITION_3D + all-reduce over peer memory, one kernel.
 * device_minmax6 (device, 6 doubles) receives [min xyz, -max xyz] over ALL ranks (DBL_MAX where no rank has points).
 * Collective: every rank of the communicator must call it the same number of times.  Returns 1 if this rank's
 * conversion produces a Vec3f64 POSITION_3D (its shard contributes), 0 if not. */
int pb200_converter_convert_into_range_with_global_bounds(pb200_converter* cv, const pb200_buffer_desc* src,
                                                          uint64_t src_begin, uint64_t src_end,
                                                          const pb200_buffer_desc* dst, uint64_t dst_begin,
                                                          uint64_t dst_end, pb200_comm* comm, double* device_minmax6);

/* ---- sharded voxel grid ------------------------------------------------------------------
 * A cloud sharded by point range over several GPUs is filtered in three steps: (1) every shard reduces its points to
 * PARTIALS on the grid of the GLOBAL bounding box (the all-reduced AABB: every shard derives the same markers,
 * voxel_grid.rs:54-79, hence the same voxel keys); (2) partials are exchanged so that each rank holds all partials of a
 * disjoint key range; (3) the rank merges them (sums added in source-rank order, counts added) and divides.  Voxel
 * keys and counts equal the single-device filter exactly; centroids within 1e-9 relative (the f64 summation order
 * differs between one in-order sum and a sum of per-shard sums).  POSITION_3D only. */
typedef struct pb200_voxel_partials pb200_voxel_partials;
typedef struct {
  uint64_t len;           /* occupied voxels */
  const uint64_t* keys;   /* device, ascending: (ix << (bits_y + bits_z)) | (iy << bits_z) | iz */
  const uint32_t* counts; /* device: points per voxel */
  const double* sums;     /* device: len * 3 position sums */
  uint32_t bits_x, bits_y, bits_z;
  uint64_t cells[3];      /* markers per axis of the global grid */
} pb200_voxel_partials_desc;
int pb200_voxelgrid_partials(pb200_ctx* ctx, const pb200_buffer_desc* src, double leaf_x, double leaf_y, double leaf_z,
                             const double global_min[3], const double global_max[3], pb200_voxel_partials** out);
/* keys / counts / sums: device arrays of m concatenated partials (source-rank order within equal keys is kept) */
int pb200_voxelgrid_merge_partials(pb200_ctx* ctx, const uint64_t* keys, const uint32_t* counts, const double* sums,
                                   uint64_t m, uint32_t bits_x, uint32_t bits_y, uint32_t bits_z,
                                   pb200_voxel_partials** out);
int pb200_voxel_partials_get(const pb200_voxel_partials* p, pb200_voxel_partials_desc* out);
/* The same for EVERY attribute of a filtered layout (voxel_grid.rs:168-689), so that a sharded cloud can be filtered with
 * all 18 built-in reductions.  Per voxel of the shard (in the order of pb200_voxel_partials_get's keys):
 *   mean attributes      one f64 column per component holding the SUM over the shard's points (Intensity, NIR, ColorRGB,
 *                        Normal); the division by the merged count and the reference's `as` cast happen after the merge
 *   max-pool attributes  one f64 column with the running maximum, starting from 0.0 like the reference (ClassificationFlags,
 *                        GpsTime, PointID)
 * and per "most common value" attribute (ReturnNumber, Classification, ...) a RUN LIST: (voxel key << 16 | value biased to
 * an unsigned 16-bit field, number of the shard's points with that value in that voxel), ascending -- per-shard modes
 * cannot be merged, per-shard histograms can.  Columns and lists are numbered in the order of dst_layout's attributes.
 * Exact for every integer attribute (u16 sums are exact in f64); f32 means differ from the single-device result only by
 * the f64 summation order (<= 1e-9 relative).  The shard must be device-resident. */
typedef struct {
  uint32_t n_columns, n_modes;
  const double* columns;                 /* device, len * n_columns, voxel-major */
  uint8_t column_is_max[64];             /* how column c merges: 0 = add, 1 = maximum */
  uint64_t mode_len[PB200_MAX_ATTRIBUTES];
  const uint64_t* mode_keys[PB200_MAX_ATTRIBUTES];   /* device, ascending */
  const uint32_t* mode_counts[PB200_MAX_ATTRIBUTES]; /* device */
} pb200_voxel_attr_partials_desc;
int pb200_voxelgrid_partials_layout(pb200_ctx* ctx, const pb200_buffer_desc* src, double leaf_x, double leaf_y, double leaf_z,
                                    const double global_min[3], const double global_max[3], const pb200_layout* dst_layout,
                                    pb200_voxel_partials** out);
int pb200_voxel_partials_get_attrs(const pb200_voxel_partials* p, pb200_voxel_attr_partials_desc* out);
/* merge of concatenated partials (`pos`: keys / counts / sums of all source ranks in rank order, len = their total;
 * `attrs`: their columns, row i belonging to pos->keys[i], and the concatenated run lists) into the filtered points of this
 * rank's key range: a library-owned COLUMNAR DEVICE buffer with dst_layout, one point per voxel in ascending key order
 * (pb200_result_buffer_voxel_keys gives the (ix, iy, iz) of each).  Mode ties resolve to the smallest value. */
int pb200_voxelgrid_merge_partials_layout(pb200_ctx* ctx, const pb200_layout* dst_layout, const pb200_voxel_partials_desc* pos,
                                          const pb200_voxel_attr_partials_desc* attrs, pb200_result_buffer** out);
/* positions_out: device, len * 3 doubles: sums / count (voxel_grid.rs:382-386) */
int pb200_voxel_partials_centroids(const pb200_voxel_partials* p, double* positions_out);
void pb200_voxel_partials_destroy(pb200_voxel_partials* p);

/* ---- kNN / radius / normals ------------------------------------------------------------------------ */
/* KdTree::nearests(q, k) for every point of the buffer against the buffer itself
 * (normal_estimation.rs:103,108): k nearest by squared distance including the point itself, ascending.
 * idx_out: len*k u32, d2_out: len*k f64 (nullable), in buf's memspace; if len < k the tail is 0xFFFFFFFF/inf. */
int pb200_knn(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint32_t* idx_out, double* d2_out);
/* The same for the points [first_query, first_query + n_queries) only (still against the whole buffer): idx_out /
 * d2_out hold n_queries * k entries, entry 0 belongs to point first_query.  This is the loop body of
 * normal_estimation.rs:106-127 for a sub-range of its `for point in points` -- the unit of work of the replicas-only
 * multi-GPU cut: every GPU holds the position column and answers its own range.  The queries are processed
 * in Morton order, so the work is proportional to n_queries.  PB200_ERR_RANGE if the range leaves the buffer. */
int pb200_knn_range(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint64_t first_query, uint64_t n_queries,
                    uint32_t* idx_out, double* d2_out);
/* all neighbours with d2 <= r*r, at most max_neighbors per point (nearest first); counts_out: len u32 */
int pb200_radius_search(pb200_ctx* ctx, const pb200_buffer_desc* buf, double radius, uint32_t max_neighbors,
                        uint32_t* idx_out, uint32_t* counts_out);
/* compute_normals, normal_estimation.rs:79-130: normals_out len*3 f64 (un-normalised, un-oriented, as the
 * reference), curvature_out len f64. PB200_ERR_TOO_FEW_POINTS if len < 3, PB200_ERR_INVALID if k < 3. */
int pb200_compute_normals(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, double* normals_out,
                          double* curvature_out);

/* compute_normals for the points [first_query, first_query + n_queries): normals_out n_queries * 3, curvature_out n_queries */
int pb200_compute_normals_range(pb200_ctx* ctx, const pb200_buffer_desc* buf, uint32_t k, uint64_t first_query, uint64_t n_queries,
                                double* normals_out, double* curvature_out);

/* ---- 3D-Tiles .pnts FeatureTable body (pasture-io/src/tiles3d) --------------------------------------------
 * The binary body is one packed array per semantic. `attrs` describes the file's layout like PntsReader::layout +
 * attribute_offsets do: name (pasture attribute name, e.g. "Position3D"), dtype, offset = byte offset of the array in
 * the body. Header and JSON stay with the caller. */
/* PntsWriter::make_compatible_layout (pnts_writer.rs:104-150) + the array offsets of create_feature_table (:234-246):
 * the supported semantics of `point_layout` (Position3D->Vec3f32, ColorRGB->Vec3u8, ColorRGBA->Vec4u8, Normal->Vec3f32,
 * everything else is dropped) in source order; *body_bytes = calc_feature_table_body_length (:300-308).
 * out_attrs must have room for 4 entries. */
int pb200_pnts_compatible_layout(const pb200_layout* point_layout, uint64_t num_points, pb200_attr* out_attrs, uint32_t* n_out,
                                 uint64_t* body_bytes);
/* PntsReader::read_into (pnts_reader.rs:294-367): points [first_point, first_point+count) of the body (same memory
 * space as dst) go to dst[0, count); file attributes the target does not have are skipped, target attributes the file
 * does not have are left alone, differing datatypes are cast. rtc_center (nullable, 3 doubles) = PntsReadPositionsMode::
 * Absolute: added to POSITION_3D of every point of dst (Vec3f32 via f64, or Vec3f64; else PB200_ERR_UNSUPPORTED, :247-283).
 * body_size = bytes readable at `body`; an array that would run past it is PB200_ERR_RANGE (UnexpectedEof in the reference). */
int pb200_pnts_read_points(pb200_ctx* ctx, const void* body, uint64_t body_size, const pb200_attr* attrs, uint32_t n_attrs,
                           uint64_t first_point, uint64_t count, const pb200_buffer_desc* dst, const double* rtc_center);
/* PntsWriter::write + write_feature_table_body (pnts_writer.rs:353-401, :310-341): all points of src, converted to
 * the compatible layout, as the FeatureTable body (zero padded arrays) at body_out (src's memory space). */
int pb200_pnts_write_points(pb200_ctx* ctx, const pb200_buffer_desc* src, void* body_out, uint64_t body_capacity);

/* ---- RANSAC segmentation (pasture-algorithms/src/segmentation.rs) ------------------------------------------
 * kind: plane = ransac_plane_{par,serial} (:180-199, :240-255), model = a,b,c,d of ax+by+cz+d=0 (4 doubles, 3 sample
 * points per model); line = ransac_line_{par,serial} (:291-310, :350-368), model = first xyz, second xyz (6 doubles,
 * 2 sample points). A point is an inlier iff distance_point_plane / distance_point_line (:31-44) < threshold,
 * evaluated in f64 exactly as written there. Requires a Vec3f64 POSITION_3D like view_attribute::<Vector3<f64>>.
 * The reference draws its samples from rand::thread_rng(); that draw is the only part that is not reproduced. */
enum { PB200_RANSAC_PLANE = 0, PB200_RANSAC_LINE = 1 };
/* generate_{plane,line}_model (:98-138) for n_models given draws: samples = n_models x 3|2 point indices (host),
 * models_out = n_models x 4|6 doubles (host), rankings_out = n_models inlier counts (host). All models are ranked in
 * one pass over the positions per 256 models. PB200_ERR_TOO_FEW_POINTS if len < 3|2, PB200_ERR_RANGE for a bad index. */
int pb200_ransac_rank_samples(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const uint64_t* samples, uint64_t n_models,
                              double distance_threshold, double* models_out, uint64_t* rankings_out);
/* the same for caller-supplied models (host) */
int pb200_ransac_rank_models(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const double* models, uint64_t n_models,
                             double distance_threshold, uint64_t* rankings_out);
/* the Vec<usize> of inlier indices of one model, ascending; indices_out (capacity entries) lives in buf's memory
 * space. *num_inliers is always set; PB200_ERR_RANGE if capacity is too small. */
int pb200_ransac_inliers(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, const double* model, double distance_threshold,
                         uint64_t* indices_out, uint64_t capacity, uint64_t* num_inliers);
/* ransac_*_serial / _par in one call: draw j is splitmix64(seed, j) % len with the reference's redraw loops, the best
 * model is the LAST one with the maximal ranking (Iterator::max_by). model_out: 4|6 doubles (host); indices_out may be
 * NULL (ranking only) or hold `capacity` >= ranking entries in buf's memory space. */
int pb200_ransac(pb200_ctx* ctx, const pb200_buffer_desc* buf, int kind, double distance_threshold, uint64_t num_of_iterations,
                 uint64_t seed, double* model_out, uint64_t* ranking_out, uint64_t* indices_out, uint64_t capacity);

/* ---- reprojection, pasture-algorithms/src/reprojection.rs:132-146,201-227 -------------------------- */
/* The reference hands two CRS strings to PROJ (proj-sys 0.22, not part of the checkout) and calls proj_trans per point.
 * PROJ strings cannot run on the GPU: the boundary takes an enumerated operation pipeline over (v0, v1, v2), either
 * composed by hand or by pb200_proj_pipeline_for_crs for the CRS pairs it knows.  Parameters (p[]) per kind:
 *   AFFINE             v := A v + b                    p[0..8] = A row-major, p[9..11] = b   (Helmert: pb200_proj_op_helmert)
 *   GEODETIC_TO_ECEF   (lat deg, lon deg, h) -> XYZ    p = a, 1/f
 *   ECEF_TO_GEODETIC   XYZ -> (lat rad, lon rad, h)    p = a, 1/f
 *   ALBERS_FWD         (lat rad, lon rad) -> (E, N)    p = a, 1/f, phi1, phi2, phi0, lam0, FE, FN (radians), Snyder 14-*
 *   SET_Z              v2 := the input z
 *   WEBMERC_FWD / _INV (lat deg, lon deg) <-> (E, N)   EPSG method 1024 (Popular Visualisation Pseudo Mercator)
 *   TMERC_FWD / _INV   (lat rad, lon rad) <-> (E, N)   p = a, 1/f, lat0, lon0 (radians), k0, FE, FN; EPSG method 9807, the
 *                                                      Krueger-series ("JHS") formulas of IOGP Guidance Note 7-2
 *   DEG2RAD_LATLON / RAD2DEG_LATLON   scale v0, v1
 * Pinning: EPSG:4326 -> EPSG:3309 by the reference's own known-answer test (reprojection.rs:275-289, 1e-4 m); the Transverse
 * Mercator, Pseudo-Mercator and Helmert operations by the worked examples of Guidance Note 7-2 and Snyder's UTM example
 * (tests/test_oracle_algorithms.py), since libproj itself is not available here.  Everything else: parity unpinned. */
enum pb200_proj_kind {
    PB200_PROJ_AFFINE = 1, PB200_PROJ_GEODETIC_TO_ECEF = 2, PB200_PROJ_ECEF_TO_GEODETIC = 3,
    PB200_PROJ_ALBERS_FWD = 4, PB200_PROJ_SET_Z = 5, PB200_PROJ_WEBMERC_FWD = 6, PB200_PROJ_TMERC_FWD = 7,
    PB200_PROJ_DEG2RAD_LATLON = 8, PB200_PROJ_RAD2DEG_LATLON = 9, PB200_PROJ_TMERC_INV = 10, PB200_PROJ_WEBMERC_INV = 11
};
typedef struct pb200_proj_op {
    uint32_t kind;
    uint32_t _pad;
    double p[12];
} pb200_proj_op;
/* fills ops (capacity >= 8) for a known CRS pair and returns the op count, or PB200_ERR_UNSUPPORTED:
 *   EPSG:4326 -> EPSG:3309;  EPSG:4326 <-> EPSG:3857;  EPSG:4326 <-> UTM on WGS 84 (EPSG:32601-32660, 32701-32760) and on
 *   ETRS89 (EPSG:25828-25838);  UTM zone <-> UTM zone.  Geographic coordinates are (lat, lon) in degrees (EPSG axis order,
 *   as in the reference's tests, reprojection.rs:266-272); z passes through. */
int pb200_proj_pipeline_for_crs(const char* source_crs, const char* target_crs, pb200_proj_op* ops, uint32_t cap);
/* a Transverse Mercator op for any zone / national grid: ellipsoid (a, 1/f), natural origin (degrees), scale factor,
 * false easting / northing; inverse != 0: (E, N) -> (lat rad, lon rad) */
int pb200_proj_op_tmerc(double a, double inv_f, double lat0_deg, double lon0_deg, double k0, double false_easting, double false_northing,
                        int inverse, pb200_proj_op* out);
/* a 7-parameter Helmert transformation of geocentric coordinates as an AFFINE op: translations in metres, rotations in
 * arc-seconds, scale difference in ppm; coordinate_frame == 0: Position Vector convention (EPSG method 1033),
 * != 0: Coordinate Frame rotation (EPSG 1032) */
int pb200_proj_op_helmert(double tx, double ty, double tz, double rx_arcsec, double ry_arcsec, double rz_arcsec, double ds_ppm,
                          int coordinate_frame, pb200_proj_op* out);
/* reproject_point_cloud_within (dst == NULL) / _between (PB200_ERR_RANGE if lengths differ) on POSITION_3D */
int pb200_reproject(pb200_ctx* ctx, const pb200_buffer_desc* src, const pb200_buffer_desc* dst_or_null,
                    const pb200_proj_op* ops, uint32_t n_ops);

/* ---- LAS point-block ingest / egress (the callers on either side of the conversion path) ---- */
typedef struct pb200_las_header {       /* the public-header fields the point path needs (LAS 1.0-1.4) */
    uint8_t version_major, version_minor, point_format, is_compressed;
    uint16_t record_length, header_size;
    uint32_t offset_to_point_data, number_of_vlrs, extra_bytes, _pad;
    uint64_t number_of_points;
    double scale[3], offset[3], min[3], max[3];
} pb200_las_header;
typedef struct pb200_las_write_stats {
    uint64_t out_of_range;              /* != 0: write_position_as_las_position would panic (write_helpers.rs:15-17) */
    uint64_t points_by_return[16];      /* [r] = points with ReturnNumber == r, r = 1..15 (raw_writers.rs:221-229,259-263) */
    int32_t has_bounds, _pad;
    double bounds_min[3], bounds_max[3];/* bounds of the written world-space positions (raw_writers.rs:28-47) */
} pb200_las_write_stats;
int pb200_las_parse_header(const void* file_bytes, uint64_t size, pb200_las_header* out);
/* RawLASReader::read_into (pasture-io/src/las/raw_readers.rs:366-383): points [first, first+count) of an uncompressed
 * LAS file image in HOST memory -> dst[dst_begin ..] in dst's own layout (default mappings of get_default_las_converter) */
int pb200_las_read_points(pb200_ctx* ctx, const void* file_bytes, uint64_t size, uint64_t first_point, uint64_t count,
                          const pb200_buffer_desc* dst, uint64_t dst_begin);
/* RawLASWriter::write_points_default_layout (raw_writers.rs:203-362): points [begin,end) of `src` (the LasPointFormatN
 * attributes by name; missing ones are written as 0) -> (end-begin) raw records of `point_format` in out_records */
int pb200_las_write_points(pb200_ctx* ctx, const pb200_buffer_desc* src, uint64_t begin, uint64_t end, int point_format,
                           const double scale[3], const double offset[3], void* out_records, int32_t out_memspace,
                           pb200_las_write_stats* stats);

/* ---- synthetic inputs (bench / test tooling; SURVEY 8d splitmix64 streams, generated in HBM) ---------- */
/* raw LAS format-0 records (20 B each) of the C2/C5 stream, points first_index .. first_index+n-1 */
int pb200_synth_las_fmt0_records(pb200_ctx* ctx, void* device_out, uint64_t first_index, uint64_t n, uint64_t seed);
/* packed Vec3f64 positions of the C3/C4 terrain stream */
int pb200_synth_terrain_positions(pb200_ctx* ctx, void* device_out, uint64_t first_index, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
