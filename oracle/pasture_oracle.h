/*
 * pasture_oracle.h -- CPU restatement of the igd-geo/pasture per-point hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (libpasture_b200.so) never links, loads or calls anything in here.
 *
 * The reference is pure Rust and cannot be built in this image (no rustc/cargo), so this is a
 * plain-C restatement that follows the reference files function by function.  Every function
 * cites the reference file:line it follows (paths relative to the reference checkout).
 *
 * Pinning (see tests/test_oracle_*.py, run under -m "not gpu"):
 *   - layout offsets/sizes     : doc-test asserts point_layout.rs:664-668,684-691,713-717,770-776,924-927,
 *                                LAS raw record sizes las_layout.rs:278, struct sizes las_types.rs:37,93
 *   - casts                    : Rust `as` language semantics (edge vectors in tests/golden/as_cast_vectors.json)
 *   - LAS read conversion      : pasture-io/resources/test/10_points_*format_N.las decoded values
 *                                (pasture-io/src/las/test_util.rs:46-183, raw_readers.rs:815-911)
 *   - LAS write transform      : round-trip property pasture-io/tests/las_io.rs:245-350
 *   - AABB                     : pasture-core/src/math/bounds.rs:291-315, test_util.rs:46-48
 *   - voxel grid               : pasture-algorithms/src/voxel_grid.rs:906-938
 *   - covariance/normals       : pasture-algorithms/src/normal_estimation.rs:503-610
 *   - reprojection             : pasture-algorithms/src/reprojection.rs:275-289 (4 KATs, 1e-4 m)
 *   - kNN                      : third-party kd-tree 0.3.0 (not vendored) -> PARITY UNPINNED on ties;
 *                                brute-force exact kNN ordered by (d2, index)
 */
#ifndef PASTURE_ORACLE_H
#define PASTURE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PointAttributeDataType, declaration order of point_layout.rs:23-68 */
enum {
    PO_U8 = 0, PO_I8 = 1, PO_U16 = 2, PO_I16 = 3, PO_U32 = 4, PO_I32 = 5, PO_U64 = 6, PO_I64 = 7,
    PO_F32 = 8, PO_F64 = 9, PO_VEC3U8 = 10, PO_VEC3U16 = 11, PO_VEC3F32 = 12, PO_VEC3I32 = 13,
    PO_VEC3F64 = 14, PO_VEC4U8 = 15, PO_BYTEARRAY = 16, PO_CUSTOM = 17
};

/* "panic" / error codes (negative) */
enum {
    PO_OK = 0,
    PO_ERR_ATTR_NOT_FOUND = -1,   /* buffer_conversion.rs:114,164,168 */
    PO_ERR_NO_CONVERSION = -2,    /* attribute_conversion.rs:267-269, buffer_conversion.rs:383-388 */
    PO_ERR_TRANSFORM_DTYPE = -3,  /* buffer_conversion.rs:209-213 */
    PO_ERR_LAYOUT_MISMATCH = -4,  /* buffer_conversion.rs:302-303 */
    PO_ERR_RANGE = -5,            /* buffer_conversion.rs:304-306 */
    PO_ERR_DUPLICATE_ATTR = -6,   /* point_layout.rs:783-788 */
    PO_ERR_OVERLAP = -7,          /* point_layout.rs:737-743 */
    PO_ERR_INVALID = -8,
    PO_ERR_TOO_FEW_POINTS = -9,   /* normal_estimation.rs:86-91,296-298 */
    PO_ERR_UNSUPPORTED = -10      /* voxel_grid.rs:452-459,682-687; raw_readers.rs:56 */
};

#define PO_MAX_ATTRS 48
#define PO_NAME_LEN 64

typedef struct {
    char name[PO_NAME_LEN];
    uint32_t dtype;
    uint64_t extra_size;   /* ByteArray length / Custom size */
    uint64_t extra_align;  /* Custom min_alignment */
    uint64_t offset;
    uint64_t size;
} po_member;

typedef struct {
    uint32_t n;
    po_member m[PO_MAX_ATTRS];
    uint64_t size;   /* memory_layout.size()  */
    uint64_t align;  /* memory_layout.align() */
} po_layout;

/* FieldAlignment (point_layout.rs:601-606): packed_n == 0 means Default */
uint64_t po_dtype_size(uint32_t dtype, uint64_t extra_size);
uint64_t po_dtype_min_alignment(uint32_t dtype, uint64_t extra_align);
void po_layout_init(po_layout* l);
int po_layout_add_attribute(po_layout* l, const char* name, uint32_t dtype, uint64_t extra_size,
                            uint64_t extra_align, uint64_t packed_n);
int po_layout_from_members_and_alignment(po_layout* l, const po_member* members, uint32_t n,
                                         uint64_t type_alignment);
int po_layout_index_by_name(const po_layout* l, const char* name);
int po_layout_index_of(const po_layout* l, const char* name, uint32_t dtype);
int po_layout_equal(const po_layout* a, const po_layout* b);

/* Casts: attribute_conversion.rs:184-343. Returns PO_ERR_NO_CONVERSION when the table has no entry. */
int po_has_conversion(uint32_t from_dtype, uint32_t to_dtype);
int po_convert_value(uint32_t from_dtype, uint32_t to_dtype, const uint8_t* from, uint8_t* to);

/* Enumerated transforms (the closures used in-tree) */
enum {
    PO_T_NONE = 0,
    PO_T_SCALE_OFFSET = 1,     /* Vec3f64: (v*s)+o ; Vec3f32: ((v as f64*s)+o) as f32   raw_readers.rs:42-55 */
    PO_T_INV_SCALE_OFFSET = 2, /* Vec3f64: (v-o)/s                                     write_helpers.rs:15-17 */
    PO_T_ADD = 3,              /* Vec3f64: v+o ; Vec3f32: (v as f64+o) as f32          pnts_reader.rs:265-277 */
    PO_T_SHIFT_MASK = 4        /* unsigned ints: (v >> shift) & mask                   raw_readers.rs:61-164 */
};
typedef struct {
    uint32_t kind;
    uint32_t shift;
    uint64_t mask;
    double s[3];
    double o[3];
} po_transform;

typedef struct {
    int32_t target_idx;
    int32_t source_idx;
    int32_t has_converter;
    int32_t has_transform;
    int32_t apply_to_source;
    int32_t _pad;
    po_transform t;
} po_mapping;

typedef struct {
    po_layout from_layout;
    po_layout to_layout;
    uint32_t n_mappings;
    po_mapping mappings[PO_MAX_ATTRS];
} po_converter;

/* buffer descriptor: VectorBuffer (AoS) or HashMapBuffer (SoA) memory, point_buffer.rs:659,1031 */
typedef struct {
    const po_layout* layout;
    int32_t columnar;       /* 0 = interleaved (aos), 1 = columnar (columns[i] per attribute in layout order) */
    uint64_t len;
    uint8_t* aos;
    uint8_t** columns;
} po_buffer;

int po_converter_for_layouts(po_converter* cv, const po_layout* from, const po_layout* to, int with_default);
int po_converter_set_custom_mapping(po_converter* cv, const char* from_name, uint32_t from_dtype,
                                    const char* to_name, uint32_t to_dtype);
int po_converter_set_custom_mapping_with_transformation(po_converter* cv, const char* from_name,
                                                        uint32_t from_dtype, const char* to_name,
                                                        uint32_t to_dtype, uint32_t transform_dtype,
                                                        const po_transform* t, int apply_to_source);
int po_convert_into_range(const po_converter* cv, const po_buffer* src, uint64_t src_begin, uint64_t src_end,
                          po_buffer* dst, uint64_t dst_begin, uint64_t dst_end);
/* multi-threaded courtesy variant: the same routine over disjoint point ranges (BASELINE.md §4 "chunked-MT") */
int po_convert_into_range_mt(const po_converter* cv, const po_buffer* src, uint64_t src_begin, uint64_t src_end,
                             po_buffer* dst, uint64_t dst_begin, uint64_t dst_end, int n_threads);

/* LAS: las_layout.rs:64-125, las_types.rs, raw_readers.rs:31-167, write_helpers.rs:10-23 */
int po_las_raw_layout(int format, po_layout* out);
int po_las_default_layout(int format, po_layout* out);
int po_las_default_converter(po_converter* cv, const po_layout* raw, const po_layout* target,
                             const double scale[3], const double offset[3]);
/* returns 0, or 1 if the point would make the reference writer panic (try_into::<i32> overflow) */
int po_las_write_position(const double world[3], const double scale[3], const double offset[3], int32_t out[3]);

/* raw_writers.rs:203-362 (default-layout writer) */
int po_las_write_points(const po_buffer* src, int format, const double scale[3], const double offset[3], uint8_t* out,
                        uint64_t counts16[16], double bmin[3], double bmax[3], uint64_t* panics);

/* bounds.rs:11-85 ; returns 1 = Some, 0 = None */
int po_calculate_bounds(const po_buffer* buf, double out_min[3], double out_max[3]);
/* minmax.rs:13-51 with T = view_dtype; returns 1 = Some, 0 = None, <0 error */
int po_minmax_attribute(const po_buffer* buf, const char* name, uint32_t attr_dtype, uint32_t view_dtype,
                        uint8_t* out_min, uint8_t* out_max);

/* math/bitmanip.rs:2-10 */
uint64_t po_expand_bits_by_3(uint64_t v);
uint64_t po_reverse_bits(uint64_t v);

/* voxel_grid.rs */
/* markers: returns count written (cap = capacity of out, returns needed count if larger) */
uint64_t po_create_markers(double bmin, double bmax, double leaf, double* out, uint64_t cap);
void po_find_leaf(const double p[3], const double* mx, uint64_t nx, const double* my, uint64_t ny,
                  const double* mz, uint64_t nz, uint64_t out[3]);
/* binary-search variant; must equal po_find_leaf (checked in tests) */
void po_find_leaf_bsearch(const double p[3], const double* mx, uint64_t nx, const double* my, uint64_t ny,
                          const double* mz, uint64_t nz, uint64_t out[3]);
/* faithful filter (sorted insert). dst must be an empty buffer description with capacity >= src->len points
 * allocated; dst->len is set to the voxel count. mode_tie_smallest: ties in centroid_most_common are
 * nondeterministic in the reference (HashMap order, voxel_grid.rs:320-328); the oracle picks the smallest value.
 * If voxel_keys != NULL it receives 3 x u64 (ix,iy,iz) per voxel in output order. */
int po_voxelgrid_filter(const po_buffer* src, double lx, double ly, double lz, po_buffer* dst,
                        uint64_t* voxel_keys, int use_sort);

/* kNN brute force (third-party kd-tree 0.3.0 `nearests`: k nearest incl. the query itself, ascending d2).
 * ties ordered by index (PARITY UNPINNED on exact ties). idx_out/d2_out: n_query x k (filled with min(k,n)) */
void po_knn_bruteforce(const double* pts, uint64_t n, const double* queries, uint64_t nq, uint32_t k,
                       uint32_t* idx_out, double* d2_out);

/* exact kNN over a hand-written kd-tree (SURVEY 8d's "nanoflann-style" CPU baseline and the checker beyond the reach of the
 * O(N^2) brute force): neighbours of the cloud's own points [q_begin, q_end), same (d2, index) order and the same d2
 * arithmetic as po_knn_bruteforce; idx_out / d2_out (n_query x k), normals_out (n_query x 3) / curv_out are each optional.
 * Tie ORDER at equal distances: parity unpinned (kd-tree 0.3.0 is not in /root/reference). */
int po_knn_kdtree(const double* pts, uint64_t n, uint64_t q_begin, uint64_t q_end, uint32_t k, uint32_t* idx_out, double* d2_out,
                  double* normals_out, double* curv_out, int threads);

/* normal_estimation.rs:198-476 on an explicit neighbourhood (k points, xyz packed) */
void po_compute_centroid(const double* pts, uint64_t k, double out[3]);
int po_compute_covariance(const double* pts, uint64_t k, double out9[9]); /* row-major 3x3 */
void po_solve_plane_parameter(const double cov9[9], double normal[3], double* curvature);
int po_normal_estimation(const double* pts, uint64_t k, double normal[3], double* curvature);
/* compute_normals (normal_estimation.rs:79-130) with brute-force kNN */
int po_compute_normals(const double* pts, uint64_t n, uint32_t k, double* normals_out, double* curv_out);

/* reprojection.rs:38-45 for the one pinned CRS pair EPSG:4326 -> EPSG:3309 (closed form) */
enum {
    PO_PROJ_AFFINE = 1,          /* p = A*v + b           params: 9 (row-major) + 3 */
    PO_PROJ_GEODETIC_TO_ECEF = 2,/* (lat_deg, lon_deg, h) -> XYZ   params: a, inv_f */
    PO_PROJ_ECEF_TO_GEODETIC = 3,/* XYZ -> (lat_rad, lon_rad, h)   params: a, inv_f */
    PO_PROJ_ALBERS_FWD = 4,      /* (lat_rad, lon_rad, h) -> (E, N, h) params: a, inv_f, phi1, phi2, phi0, lam0, x0, y0 (radians) */
    PO_PROJ_SET_Z = 5,           /* z := saved input z (z passthrough)  */
    PO_PROJ_WEBMERC_FWD = 6,     /* (lat_deg, lon_deg, h) -> EPSG:3857 */
    PO_PROJ_TMERC_FWD = 7,       /* (lat_rad, lon_rad, h) -> (E,N,h) params: a, inv_f, lat0, lon0, k0, FE, FN; GN7-2 JHS formulas */
    PO_PROJ_DEG2RAD_LATLON = 8, PO_PROJ_RAD2DEG_LATLON = 9,
    PO_PROJ_TMERC_INV = 10,      /* (E,N,h) -> (lat_rad, lon_rad, h), same params */
    PO_PROJ_WEBMERC_INV = 11     /* EPSG:3857 -> (lat_deg, lon_deg, h) */
};
typedef struct {
    uint32_t kind;
    uint32_t _pad;
    double p[12];
} po_proj_op;
void po_reproject(const po_proj_op* ops, uint32_t n_ops, const double* in_xyz, double* out_xyz, uint64_t n);
/* fills ops[] (cap >= 8) with the EPSG:4326 -> EPSG:3309 pipeline; returns n_ops */
uint32_t po_pipeline_epsg4326_to_3309(po_proj_op* ops);

/* deterministic synthetic generators, SURVEY 8(d) */
uint64_t po_splitmix64(uint64_t seed, uint64_t j);

/* ---- RANSAC segmentation (pasture-algorithms/src/segmentation.rs) --------------------------------------
 * The reference draws its sample indices from rand::thread_rng() (:49-59, :82-89), so WHICH models are tried is not
 * reproducible there; what is deterministic -- and restated here -- is everything after the draw: the model from the
 * sampled points (:60-76, :91-95), the point/model distance (:31-44), the strict `<` inlier test with the ranking
 * and index list (:98-138) and the choice of the best model (max_by ranking: the LAST maximum, :192-199).
 * kind 0 = plane (3 samples per model, model = a,b,c,d), kind 1 = line (2 samples, model = first xyz, second xyz). */
void po_ransac_model_from_samples(int kind, const double* pts, const uint64_t* samples, double* model);
double po_ransac_distance(int kind, const double* model, const double p[3]);
/* rankings[h] = number of points with distance < threshold for model h (models: 4 or 6 doubles each) */
void po_ransac_rank_models(int kind, const double* pts, uint64_t n, const double* models, uint64_t n_models,
                           double threshold, uint64_t* rankings);
/* indices of the inliers of one model, ascending; returns their number */
uint64_t po_ransac_inliers(int kind, const double* pts, uint64_t n, const double* model, double threshold, uint64_t* indices);
/* the draw used by the GPU library's convenience entry point: draw j = splitmix64(seed, j) % n, redrawn like the
 * reference's `while` loops until the 3 (2) indices differ. Fills samples (n_models x 3|2). */
void po_ransac_draw_samples(int kind, uint64_t n, uint64_t n_models, uint64_t seed, uint64_t* samples);
void po_gen_las_fmt0_records(uint8_t* out20, uint64_t first_index, uint64_t n, uint64_t seed);
void po_gen_c1_points(uint8_t* out35, uint64_t first_index, uint64_t n, uint64_t seed, const double offset[3]);
void po_gen_terrain_positions(double* out_xyz, uint64_t first_index, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
