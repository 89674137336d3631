/* abi_consumer.c -- a consumer of include/pasture_b200.h that is neither Python nor C++: plain C11, compiled by gcc.
 *
 * There is no Rust toolchain in this image, so bindings/rust/pasture_b200_sys.rs cannot be compiled.  This program is the
 * closest available proof that the header is a usable C ABI: (1) the header compiles as C, (2) the struct layouts are the
 * ones the ctypes mirror (pasture_b200/_lib.py) and the #[repr(C)] declarations of pasture_b200_sys.rs assume
 * (static asserts + printed table, cross-checked by tests/test_abi_consumer.py), (3) one conversion through the
 * reference-shaped call sequence -- PointLayout x2 -> get_default_las_converter -> convert_into_range
 * (buffer_conversion.rs:292-359) on HOST buffers -- decodes the ten points of the reference's LAS format-0 fixture to
 * the values pasture-io/src/las/test_util.rs:46-130 expects.
 *
 *   abi_consumer layout     print sizeof/offsetof lines (no GPU needed)
 *   abi_consumer convert    run the conversion; exit 0 = values match, 77 = no CUDA device (PB200_ERR_NO_DEVICE), 1 = mismatch
 */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "pasture_b200.h"

_Static_assert(sizeof(pb200_attr) == 104, "pb200_attr");
_Static_assert(offsetof(pb200_attr, dtype) == 64 && offsetof(pb200_attr, extra_size) == 72 && offsetof(pb200_attr, offset) == 88 &&
               offsetof(pb200_attr, size) == 96, "pb200_attr fields");
_Static_assert(sizeof(pb200_buffer_desc) == 40, "pb200_buffer_desc");
_Static_assert(offsetof(pb200_buffer_desc, kind) == 8 && offsetof(pb200_buffer_desc, memspace) == 12 && offsetof(pb200_buffer_desc, len) == 16 &&
               offsetof(pb200_buffer_desc, aos) == 24 && offsetof(pb200_buffer_desc, columns) == 32, "pb200_buffer_desc fields");
_Static_assert(sizeof(pb200_transform) == 64, "pb200_transform");
_Static_assert(offsetof(pb200_transform, mask) == 8 && offsetof(pb200_transform, s) == 16 && offsetof(pb200_transform, o) == 40, "pb200_transform fields");
_Static_assert(sizeof(pb200_proj_op) == 104, "pb200_proj_op");
_Static_assert(sizeof(pb200_las_header) == 128, "pb200_las_header");
_Static_assert(offsetof(pb200_las_header, number_of_points) == 24 && offsetof(pb200_las_header, scale) == 32, "pb200_las_header fields");
_Static_assert(sizeof(pb200_las_write_stats) == 192, "pb200_las_write_stats");
_Static_assert(sizeof(pb200_voxel_partials_desc) == 72, "pb200_voxel_partials_desc");

static int print_layout(void) {
#define SZ(T) printf("sizeof %s %zu\n", #T, sizeof(T))
#define OFF(T, f) printf("offsetof %s %s %zu\n", #T, #f, offsetof(T, f))
    SZ(pb200_attr); OFF(pb200_attr, name); OFF(pb200_attr, dtype); OFF(pb200_attr, extra_size); OFF(pb200_attr, extra_align);
    OFF(pb200_attr, offset); OFF(pb200_attr, size);
    SZ(pb200_buffer_desc); OFF(pb200_buffer_desc, layout); OFF(pb200_buffer_desc, kind); OFF(pb200_buffer_desc, memspace);
    OFF(pb200_buffer_desc, len); OFF(pb200_buffer_desc, aos); OFF(pb200_buffer_desc, columns);
    SZ(pb200_transform); OFF(pb200_transform, kind); OFF(pb200_transform, shift); OFF(pb200_transform, mask); OFF(pb200_transform, s);
    OFF(pb200_transform, o);
    SZ(pb200_proj_op); OFF(pb200_proj_op, kind); OFF(pb200_proj_op, p);
    SZ(pb200_las_header); OFF(pb200_las_header, point_format); OFF(pb200_las_header, record_length); OFF(pb200_las_header, offset_to_point_data);
    OFF(pb200_las_header, number_of_points); OFF(pb200_las_header, scale); OFF(pb200_las_header, offset); OFF(pb200_las_header, min);
    OFF(pb200_las_header, max);
    SZ(pb200_las_write_stats); OFF(pb200_las_write_stats, points_by_return); OFF(pb200_las_write_stats, has_bounds);
    OFF(pb200_las_write_stats, bounds_min); OFF(pb200_las_write_stats, bounds_max);
    SZ(pb200_voxel_partials_desc); OFF(pb200_voxel_partials_desc, keys); OFF(pb200_voxel_partials_desc, counts); OFF(pb200_voxel_partials_desc, sums);
    OFF(pb200_voxel_partials_desc, bits_x); OFF(pb200_voxel_partials_desc, cells);
    printf("abi_version %d\n", pb200_abi_version());
    return 0;
}

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ < 0) {                                                                   \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pb200_last_error());           \
            return rc_ == PB200_ERR_NO_DEVICE ? 77 : 1;                                  \
        }                                                                                \
    } while (0)

static int run_convert(void) {
    /* host-side layout logic needs no device: raw LAS format 0 is 20 packed bytes, the default layout 35 (las_types.rs:37) */
    pb200_layout *raw = NULL, *def = NULL;
    CHECK(pb200_las_raw_layout(0, &raw));
    CHECK(pb200_las_default_layout(0, &def));
    if (pb200_layout_size_of_point_entry(raw) != 20 || pb200_layout_size_of_point_entry(def) != 35 || pb200_layout_num_attributes(def) != 10) {
        fprintf(stderr, "unexpected LAS format-0 layouts\n");
        return 1;
    }
    /* the ten records of pasture-io/resources/test/10_points_format_0.las: X = Y = Z = i, intensity 255 i, flags bytes
     * 00 c9 12 db 24 ed 36 ff 00 c9, classification = scan angle rank = user data = point source id = i */
    static const unsigned char flags[10] = {0x00, 0xc9, 0x12, 0xdb, 0x24, 0xed, 0x36, 0xff, 0x00, 0xc9};
    unsigned char records[10 * 20];
    for (int i = 0; i < 10; ++i) {
        unsigned char* r = records + 20 * i;
        int32_t xyz[3] = {i, i, i};
        uint16_t inten = (uint16_t)(255 * i), psid = (uint16_t)i;
        memcpy(r, xyz, 12);
        memcpy(r + 12, &inten, 2);
        r[14] = flags[i];
        r[15] = (unsigned char)i;
        r[16] = (unsigned char)i;
        r[17] = (unsigned char)i;
        memcpy(r + 18, &psid, 2);
    }
    pb200_ctx* ctx = NULL;
    CHECK(pb200_ctx_create(0, &ctx));
    const double scale[3] = {1.0, 1.0, 1.0}, offset[3] = {0.0, 0.0, 0.0};
    pb200_converter* cv = NULL;
    CHECK(pb200_las_default_converter(ctx, raw, def, scale, offset, &cv));
    double pos[30];
    uint16_t intensity[10], psid[10];
    uint8_t rn[10], nr[10], sdf[10], eof[10], cls[10], ud[10];
    int8_t sar[10];
    void* columns[10] = {pos, intensity, rn, nr, sdf, eof, cls, sar, ud, psid}; /* LasPointFormat0 attribute order, las_types.rs:10-37 */
    pb200_buffer_desc src = {raw, PB200_INTERLEAVED, PB200_HOST, 10, records, NULL};
    pb200_buffer_desc dst = {def, PB200_COLUMNAR, PB200_HOST, 10, NULL, columns};
    CHECK(pb200_converter_convert_into_range(cv, &src, 0, 10, &dst, 0, 10, NULL));
    static const uint8_t want_rn[10] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1}; /* test_util.rs:80-82, 88-90 */
    int bad = 0;
    for (int i = 0; i < 10; ++i) {
        bad += pos[3 * i] != (double)i || pos[3 * i + 1] != (double)i || pos[3 * i + 2] != (double)i; /* :50-63 */
        bad += intensity[i] != 255 * i;                                                                 /* :65-78 */
        bad += rn[i] != want_rn[i] || nr[i] != want_rn[i];
        bad += sdf[i] != (i & 1) || eof[i] != (i & 1);                                                  /* :104-110 */
        bad += cls[i] != i || sar[i] != i || ud[i] != i || psid[i] != i;                                /* :112-130 */
    }
    /* the same call with a mismatching layout must fail like the reference's assert (buffer_conversion.rs:302-303) */
    pb200_buffer_desc wrong = {def, PB200_INTERLEAVED, PB200_HOST, 10, records, NULL};
    if (pb200_converter_convert_into_range(cv, &wrong, 0, 10, &dst, 0, 10, NULL) != PB200_ERR_LAYOUT_MISMATCH) ++bad;
    double bmin[3], bmax[3];
    int some = 0;
    CHECK(pb200_calculate_bounds(ctx, &dst, bmin, bmax, &some)); /* bounds 0..9, test_util.rs:46-48 */
    bad += !some || bmin[0] != 0.0 || bmax[2] != 9.0;
    pb200_converter_destroy(cv);
    pb200_ctx_destroy(ctx);
    pb200_layout_destroy(raw);
    pb200_layout_destroy(def);
    printf("abi_consumer convert: %s (kernel launches: %llu)\n", bad ? "MISMATCH" : "ok", (unsigned long long)pb200_kernel_launch_count());
    return bad ? 1 : 0;
}

/* host side only (no device): a planning-only converter (ctx == NULL) for the read path of the LAS reader and the tile
 * schedule a 1 M-point conversion from interleaved records into columns would run */
static int run_plan(void) {
    pb200_layout *raw = NULL, *def = NULL;
    CHECK(pb200_las_raw_layout(0, &raw));
    CHECK(pb200_las_default_layout(0, &def));
    const double scale[3] = {0.001, 0.001, 0.001}, offset[3] = {500000.0, 5400000.0, 100.0};
    pb200_converter* cv = NULL;
    CHECK(pb200_las_default_converter(NULL, raw, def, scale, offset, &cv));
    void* columns[10];
    for (int i = 0; i < 10; ++i) columns[i] = (void*)(uintptr_t)(0x100000000ull + 0x10000000ull * (unsigned)i); /* addresses only */
    pb200_buffer_desc src = {raw, PB200_INTERLEAVED, PB200_DEVICE, 1u << 20, (void*)(uintptr_t)0x80000000ull, NULL};
    pb200_buffer_desc dst = {def, PB200_COLUMNAR, PB200_DEVICE, 1u << 20, NULL, columns};
    static char text[1 << 16];
    int n = pb200_converter_describe_schedule(cv, &src, 0, 1u << 20, &dst, 0, 0, text, sizeof text);
    if (n <= 0) {
        fprintf(stderr, "describe_schedule -> %d (%s)\n", n, pb200_last_error());
        return 1;
    }
    fputs(text, stdout);
    /* and conversions with it are refused */
    int rc = pb200_converter_convert_into_range(cv, &src, 0, 16, &dst, 0, 16, NULL);
    pb200_converter_destroy(cv);
    pb200_layout_destroy(raw);
    pb200_layout_destroy(def);
    return rc == PB200_ERR_NO_DEVICE ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "plan") == 0) return run_plan();
    if (argc > 1 && strcmp(argv[1], "layout") == 0) return print_layout();
    if (argc > 1 && strcmp(argv[1], "convert") == 0) return run_convert();
    fprintf(stderr, "usage: abi_consumer layout|plan|convert\n");
    return 2;
}
