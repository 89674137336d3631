// pnts.cu -- the per-point part of the 3D-Tiles .pnts reader and writer (pasture-io/src/tiles3d/pnts_reader.rs:294-367,
// :247-283 and pnts_writer.rs:104-150, :234-266, :353-401).  The FeatureTable binary body is a columnar block
// (one tightly packed array per semantic at a byte offset from the JSON header); reading it is a columnar -> any
// conversion with the RTC_CENTER addition as the target-side transform, writing it is an any -> columnar conversion to
// the default semantic datatypes.  Header / JSON handling stays on the host side of the caller.
#include "internal.h"

namespace pb200 {

struct LayoutHolder {  // RAII over the C API handle
    pb200_layout* l = nullptr;
    ~LayoutHolder() { if (l) pb200_layout_destroy(l); }
};
struct ConverterHolder {
    pb200_converter* c = nullptr;
    ~ConverterHolder() { if (c) pb200_converter_destroy(c); }
};

static uint64_t align8(uint64_t v) { return (v + 7) & ~7ull; }

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_pnts_compatible_layout(const pb200_layout* point_layout, uint64_t num_points, pb200_attr* out_attrs, uint32_t* n_out,
                                 uint64_t* body_bytes) {
    if (!point_layout || !out_attrs || !n_out) return set_error(PB200_ERR_INVALID, "null argument");
    uint32_t k = 0;
    uint64_t off = 0;
    for (const pb200_attr& a : point_layout->attrs) {  // pnts_writer.rs:118-148, source attribute order
        uint32_t dt;
        if (!strcmp(a.name, "Position3D")) dt = PB200_VEC3F32;
        else if (!strcmp(a.name, "ColorRGB")) dt = PB200_VEC3U8;
        else if (!strcmp(a.name, "ColorRGBA")) dt = PB200_VEC4U8;
        else if (!strcmp(a.name, "Normal")) dt = PB200_VEC3F32;
        else continue;
        pb200_attr o{};
        strncpy(o.name, a.name, PB200_MAX_NAME - 1);
        o.dtype = dt;
        o.size = pb200_dtype_size(dt, 0);
        o.offset = off;  // create_feature_table :236-246: every array is padded to 8 bytes
        off += align8(o.size * num_points);
        out_attrs[k++] = o;
    }
    *n_out = k;
    if (body_bytes) *body_bytes = off;  // calc_feature_table_body_length :300-308
    return PB200_OK;
}

static int body_desc(const pb200_attr* attrs, uint32_t n_attrs, const void* body, int memspace, uint64_t first, uint64_t count,
                     LayoutHolder* L, std::vector<void*>* cols, pb200_buffer_desc* d) {
    PB_TRY(pb200_layout_create(&L->l));
    cols->clear();
    for (uint32_t i = 0; i < n_attrs; ++i) {
        PB_TRY(pb200_layout_add_attribute(L->l, attrs[i].name, attrs[i].dtype, attrs[i].extra_size, attrs[i].extra_align, 1));
        const uint64_t sz = pb200_dtype_size(attrs[i].dtype, attrs[i].extra_size);
        cols->push_back((uint8_t*)body + attrs[i].offset + first * sz);  // pnts_reader.rs:313-316
    }
    d->layout = L->l;
    d->kind = PB200_COLUMNAR;
    d->memspace = memspace;
    d->len = count;
    d->aos = nullptr;
    d->columns = cols->data();
    return PB200_OK;
}

int pb200_pnts_read_points(pb200_ctx* ctx, const void* body, uint64_t body_size, const pb200_attr* attrs, uint32_t n_attrs,
                           uint64_t first_point, uint64_t count, const pb200_buffer_desc* dst, const double* rtc_center) {
    if (!ctx || !attrs || (!body && count)) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(dst, "point buffer"));
    if (n_attrs > PB200_MAX_ATTRIBUTES) return set_error(PB200_ERR_INVALID, "too many attributes");
    // offsets and POINTS_LENGTH come from an untrusted JSON header: every array must lie inside the body (the reference's
    // reader fails with UnexpectedEof when it runs off the file, pnts_reader.rs:313-345); overflow-safe arithmetic
    for (uint32_t i = 0; i < n_attrs && count; ++i) {
        const uint64_t sz = pb200_dtype_size(attrs[i].dtype, attrs[i].extra_size);
        const uint64_t off = attrs[i].offset;
        if (first_point > UINT64_MAX - count) return set_error(PB200_ERR_RANGE, "point range overflows");
        const uint64_t last = first_point + count;
        if (off > body_size || (sz && last > (body_size - off) / sz))
            return set_error(PB200_ERR_RANGE, "unexpected end of file: array of attribute %s (offset %llu, %llu points of %llu bytes) "
                             "runs past the %llu-byte FeatureTable image", attrs[i].name, (unsigned long long)off,
                             (unsigned long long)last, (unsigned long long)sz, (unsigned long long)body_size);
    }
    if (count > dst->len) return set_error(PB200_ERR_RANGE, "point buffer holds %llu points, %llu requested", (unsigned long long)dst->len, (unsigned long long)count);
    // RTC_CENTER goes onto POSITION_3D of the WHOLE target buffer, whatever was read (pnts_reader.rs:247-283, :360-363)
    const int pos_t = pb200_layout_index_by_name(dst->layout, "Position3D");
    uint32_t pos_dtype = 0;
    pb200_transform add{};
    const bool rtc = rtc_center && pos_t >= 0;
    if (rtc) {
        pos_dtype = dst->layout->attrs[(size_t)pos_t].dtype;
        if (pos_dtype != PB200_VEC3F32 && pos_dtype != PB200_VEC3F64)  // :280
            return set_error(PB200_ERR_UNSUPPORTED, "Unsupported datatype for POSITION_3D attribute");
        add.kind = PB200_T_ADD;
        for (int c = 0; c < 3; ++c) { add.s[c] = 1.0; add.o[c] = rtc_center[c]; }
    }
    bool fused = false;
    if (count) {
        LayoutHolder L;
        std::vector<void*> cols;
        pb200_buffer_desc src;
        PB_TRY(body_desc(attrs, n_attrs, body, dst->memspace, first_point, count, &L, &cols, &src));
        ConverterHolder cv;
        // every file attribute that the target layout has by name, through get_converter_for_attributes (:307-327)
        PB_TRY(pb200_converter_create(ctx, L.l, dst->layout, 1, &cv.c));
        const int pos_s = pb200_layout_index_by_name(L.l, "Position3D");
        if (rtc && pos_s >= 0) {
            PB_TRY(pb200_converter_set_custom_mapping_with_transformation(cv.c, "Position3D", L.l->attrs[(size_t)pos_s].dtype, "Position3D",
                                                                          pos_dtype, pos_dtype, &add, 0));
            fused = true;
        }
        if (pb200_converter_num_mappings(cv.c) > 0)
            PB_TRY(pb200_converter_convert_into_range(cv.c, &src, 0, count, dst, 0, count, nullptr));
        else
            fused = false;
    }
    if (rtc) {
        const uint64_t begin = fused ? count : 0;
        if (begin < dst->len) {  // the points of the buffer that this read did not produce
            pb200_buffer_desc rest = *dst;
            std::vector<void*> rcols;
            rest.len = dst->len - begin;
            if (dst->kind == PB200_INTERLEAVED) rest.aos = (uint8_t*)dst->aos + begin * dst->layout->size;
            else {
                for (size_t a = 0; a < dst->layout->attrs.size(); ++a) rcols.push_back((uint8_t*)dst->columns[a] + begin * dst->layout->attrs[a].size);
                rest.columns = rcols.data();
            }
            PB_TRY(pb200_transform_attribute(ctx, &rest, "Position3D", pos_dtype, &add));
        }
    }
    return PB200_OK;
}

int pb200_pnts_write_points(pb200_ctx* ctx, const pb200_buffer_desc* src, void* body_out, uint64_t body_capacity) {
    if (!ctx) return set_error(PB200_ERR_INVALID, "null argument");
    PB_TRY(validate_desc(src, "point buffer"));
    pb200_attr attrs[PB200_MAX_ATTRIBUTES];
    uint32_t n_attrs = 0;
    uint64_t bytes = 0;
    PB_TRY(pb200_pnts_compatible_layout(src->layout, src->len, attrs, &n_attrs, &bytes));
    if (bytes > body_capacity) return set_error(PB200_ERR_RANGE, "FeatureTable body needs %llu bytes", (unsigned long long)bytes);
    if (bytes == 0) return PB200_OK;
    if (!body_out) return set_error(PB200_ERR_INVALID, "null argument");
    LayoutHolder L;
    std::vector<void*> cols;
    pb200_buffer_desc dst;
    PB_TRY(body_desc(attrs, n_attrs, body_out, src->memspace, 0, src->len, &L, &cols, &dst));
    ConverterHolder cv;
    PB_TRY(pb200_converter_create(ctx, src->layout, L.l, 0, &cv.c));  // attribute_converters, pnts_writer.rs:130-146
    if (src->len) PB_TRY(pb200_converter_convert_into_range(cv.c, src, 0, src->len, &dst, 0, src->len, nullptr));
    for (uint32_t i = 0; i < n_attrs; ++i) {  // zero padding after every array, write_feature_table_body :319-327
        const uint64_t used = attrs[i].size * src->len, pad = align8(used) - used;
        if (!pad) continue;
        uint8_t* p = (uint8_t*)body_out + attrs[i].offset + used;
        if (src->memspace == PB200_HOST) memset(p, 0, (size_t)pad);
        else PB_TRY(pb200_memset_device(ctx, p, 0, pad));
    }
    return PB200_OK;
}

}  // extern "C"
