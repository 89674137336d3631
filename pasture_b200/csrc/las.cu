// las.cu -- LAS point-block ingest / egress on the device: the steps on either side of the conversion
// path.  File handling stays on the host and is minimal (fixed-width public header, pasture-io/src/las/raw_readers.rs:
// 191-241 uses the `las` crate for it); the per-point work runs through the conversion kernel:
//   read : raw records (point format 0-10, optional extra bytes) -> any target layout, exactly what
//          RawLASReader::read_into does per chunk (raw_readers.rs:299-352) with get_default_las_converter (:31-167)
//   write: LasPointFormatN default layout -> raw records, RawLASWriter::write_points_default_layout
//          (raw_writers.rs:203-362): truncating (p-offset)/scale positions (write_helpers.rs:10-23), bit-field
//          packing (:26-51), points-by-return counts (:221-229,259-263) and the running bounds (:28-47)
#include "internal.h"

namespace pb200 {

static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static double rdf64(const uint8_t* p) { double v; memcpy(&v, p, 8); return v; }

static int raw_layout_with_extra_bytes(int format, uint32_t extra, pb200_layout** out) {
    PB_TRY(pb200_las_raw_layout(format, out));
    if (extra) {  // las_layout.rs:171-181 (undescribed extra bytes as one raw byte array, packed(1))
        int rc = pb200_layout_add_attribute(*out, "UndescribedExtraBytes", PB200_BYTEARRAY, extra, 0, 1);
        if (rc < 0) { pb200_layout_destroy(*out); *out = nullptr; return rc; }
    }
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_las_parse_header(const void* file_bytes, uint64_t size, pb200_las_header* out) {
    if (!file_bytes || !out) return set_error(PB200_ERR_INVALID, "null argument");
    const uint8_t* b = (const uint8_t*)file_bytes;
    if (size < 227 || memcmp(b, "LASF", 4) != 0) return set_error(PB200_ERR_INVALID, "not a LAS file (missing LASF signature / truncated header)");
    memset(out, 0, sizeof(*out));
    out->version_major = b[24];
    out->version_minor = b[25];
    out->header_size = rd16(b + 94);
    out->offset_to_point_data = rd32(b + 96);
    out->number_of_vlrs = rd32(b + 100);
    const uint8_t fmt = b[104];
    out->is_compressed = (fmt & 0xC0) ? 1 : 0;
    out->point_format = fmt & 0x3F;
    out->record_length = rd16(b + 105);
    out->number_of_points = rd32(b + 107);
    for (int c = 0; c < 3; ++c) { out->scale[c] = rdf64(b + 131 + 8 * c); out->offset[c] = rdf64(b + 155 + 8 * c); }
    out->max[0] = rdf64(b + 179); out->min[0] = rdf64(b + 187);
    out->max[1] = rdf64(b + 195); out->min[1] = rdf64(b + 203);
    out->max[2] = rdf64(b + 211); out->min[2] = rdf64(b + 219);
    if (out->version_major == 1 && out->version_minor >= 4 && out->header_size >= 375 && size >= 375) {
        const uint64_t n64 = rd64(b + 247);
        if (out->number_of_points == 0) out->number_of_points = n64;
    }
    if (out->point_format > 10) return set_error(PB200_ERR_UNSUPPORTED, "Unsupported LAS point format %u", out->point_format);
    pb200_layout* raw = nullptr;
    PB_TRY(pb200_las_raw_layout(out->point_format, &raw));
    const uint64_t base = raw->size;
    pb200_layout_destroy(raw);
    if (out->record_length < base) return set_error(PB200_ERR_INVALID, "point record length %u is smaller than point format %u", out->record_length, out->point_format);
    out->extra_bytes = (uint32_t)(out->record_length - base);
    return PB200_OK;
}

int pb200_las_read_points(pb200_ctx* ctx, const void* file_bytes, uint64_t size, uint64_t first_point, uint64_t count,
                          const pb200_buffer_desc* dst, uint64_t dst_begin) {
    if (!ctx) return set_error(PB200_ERR_INVALID, "null context");
    pb200_las_header h;
    PB_TRY(pb200_las_parse_header(file_bytes, size, &h));
    if (h.is_compressed) return set_error(PB200_ERR_UNSUPPORTED, "LAZ-compressed point data is out of scope (decompress on the host first)");
    PB_TRY(validate_desc(dst, "target buffer"));
    if (first_point > h.number_of_points || count > h.number_of_points - first_point)
        return set_error(PB200_ERR_RANGE, "requested points %llu..%llu but the file holds %llu", (unsigned long long)first_point,
                         (unsigned long long)(first_point + count), (unsigned long long)h.number_of_points);
    // number_of_points is a 64-bit field of an untrusted header: no products that can wrap
    if (h.record_length == 0 || (uint64_t)h.offset_to_point_data > size ||
        h.number_of_points > (size - (uint64_t)h.offset_to_point_data) / h.record_length)
        return set_error(PB200_ERR_RANGE, "file is truncated: point block exceeds the buffer");
    if (dst_begin > dst->len || count > dst->len - dst_begin) return set_error(PB200_ERR_RANGE, "point_buffer.len() must be >= count");
    if (dst->len < dst_begin + count) return set_error(PB200_ERR_RANGE, "point_buffer.len() must be >= count");  // raw_readers.rs:374-376
    pb200_layout* raw = nullptr;
    PB_TRY(raw_layout_with_extra_bytes(h.point_format, h.extra_bytes, &raw));
    pb200_converter* cv = nullptr;
    int rc = pb200_las_default_converter(ctx, raw, dst->layout, h.scale, h.offset, &cv);
    if (rc == PB200_OK) {
        pb200_buffer_desc src;
        src.layout = raw;
        src.kind = PB200_INTERLEAVED;
        src.memspace = PB200_HOST;
        src.len = h.number_of_points;
        src.aos = (void*)((const uint8_t*)file_bytes + h.offset_to_point_data);
        src.columns = nullptr;
        rc = pb200_converter_convert_into_range(cv, &src, first_point, first_point + count, dst, dst_begin, dst_begin + count, nullptr);
        if (rc == PB200_OK && dst->memspace == PB200_DEVICE) cudaStreamSynchronize(ctx->stream);  // staging buffers die with cv
    }
    pb200_converter_destroy(cv);
    pb200_layout_destroy(raw);
    return rc;
}

int pb200_las_write_points(pb200_ctx* ctx, const pb200_buffer_desc* src, uint64_t begin, uint64_t end, int point_format,
                           const double scale[3], const double offset[3], void* out_records, int32_t out_memspace,
                           pb200_las_write_stats* stats) {
    if (!ctx || !scale || !offset || !stats) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(src, "source buffer"));
    if (end < begin || end > src->len) return set_error(PB200_ERR_RANGE, "point range out of bounds");
    for (int c = 0; c < 3; ++c)
        if (scale[c] == 0.0) return set_error(PB200_ERR_INVALID, "LAS scale factors must not be zero");  // raw_writers.rs:143-148
    memset(stats, 0, sizeof(*stats));
    const uint64_t n = end - begin;
    if (n == 0) return PB200_OK;
    if (!out_records) return set_error(PB200_ERR_INVALID, "null output");
    PB_DEVICE(ctx);
    pb200_layout* raw = nullptr;
    PB_TRY(pb200_las_raw_layout(point_format, &raw));
    const bool extended = point_format >= 6;
    pb200_converter* cv = nullptr;
    int rc = pb200_converter_create(ctx, src->layout, raw, 1, &cv);
    auto done = [&](int r) { pb200_converter_destroy(cv); pb200_layout_destroy(raw); return r; };
    if (rc < 0) return done(rc);
    const int pi = pb200_layout_index_of(src->layout, "Position3D", PB200_VEC3F64);
    if (pi < 0) return done(set_error(PB200_ERR_ATTR_NOT_FOUND, "the LAS writer needs a Vec3f64 Position3D attribute"));
    pb200_transform t{};
    t.kind = PB200_T_INV_SCALE_OFFSET;
    for (int c = 0; c < 3; ++c) { t.s[c] = scale[c]; t.o[c] = offset[c]; }
    rc = pb200_converter_set_custom_mapping_with_transformation(cv, "Position3D", PB200_VEC3F64, "LASLocalPosition", PB200_VEC3I32,
                                                                PB200_VEC3F64, &t, 1);
    if (rc < 0) return done(rc);
    {  // write_helpers.rs:26-51
        std::vector<const char*> names;
        std::vector<uint32_t> masks, shifts;
        auto add = [&](const char* nm, uint32_t mask, uint32_t shift) {
            if (pb200_layout_index_of(src->layout, nm, PB200_U8) >= 0) { names.push_back(nm); masks.push_back(mask); shifts.push_back(shift); }
        };
        if (extended) {
            add("ReturnNumber", 0xF, 0); add("NumberOfReturns", 0xF, 4); add("ClassificationFlags", 0xF, 8);
            add("ScannerChannel", 0x3, 12); add("ScanDirectionFlag", 0x1, 14); add("EdgeOfFlightLine", 0x1, 15);
        } else {
            add("ReturnNumber", 0x7, 0); add("NumberOfReturns", 0x7, 3); add("ScanDirectionFlag", 0x1, 6); add("EdgeOfFlightLine", 0x1, 7);
        }
        if (!names.empty()) {
            rc = pb200_converter_set_packed_mapping(cv, extended ? "LASExtendedFlags" : "LASBasicFlags", extended ? PB200_U16 : PB200_U8,
                                                    (uint32_t)names.size(), names.data(), masks.data(), shifts.data());
            if (rc < 0) return done(rc);
        }
    }
    pb200_buffer_desc dst;
    dst.layout = raw;
    dst.kind = PB200_INTERLEAVED;
    dst.memspace = out_memspace;
    dst.len = n;
    dst.aos = out_records;
    dst.columns = nullptr;
    // one pass: records (every byte written: attributes the source lacks become 0), out-of-range count, points by return,
    // bounds of the written positions
    const int ri = pb200_layout_index_of(src->layout, "ReturnNumber", PB200_U8);
    EgressStats es;
    rc = convert_range_egress(cv, src, begin, end, &dst, 0, n, pi, ri, &es);
    if (rc < 0) return done(rc);
    stats->out_of_range = es.out_of_range;
    for (int b = 1; b < 16; ++b) stats->points_by_return[b] = es.hist[b];
    if (es.bounds_tracked) {
        if (!es.has_bounds)  // calculate_bounds on all-NaN positions: AABB::from_min_max panics (bounds.rs:21-26)
            return done(set_error(PB200_ERR_INVALID, "AABB::from_min_max: Minimum position must be <= maximum position!"));
        stats->has_bounds = 1;
        for (int c = 0; c < 3; ++c) { stats->bounds_min[c] = es.src_min[c]; stats->bounds_max[c] = es.src_max[c]; }
    }
    if (out_memspace == PB200_DEVICE) cudaStreamSynchronize(ctx->stream);
    return done(PB200_OK);
}

}  // extern "C"
