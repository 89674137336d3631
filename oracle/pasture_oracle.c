/*
 * pasture_oracle.c -- CPU restatement of the igd-geo/pasture per-point hot path (TEST INFRASTRUCTURE).
 * See pasture_oracle.h for the rules on who may load this and for the pinning status.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: Rust never contracts a*b+c into an FMA)
 *
 * Reference paths are relative to the reference checkout (igd-geo/pasture v0.5.0 @ 1b0b39c).
 */
#include "pasture_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * L: data types and PointLayout  (pasture-core/src/layout/point_layout.rs)
 * ---------------------------------------------------------------------------------------------- */

/* point_layout.rs:72-97 */
uint64_t po_dtype_size(uint32_t dtype, uint64_t extra_size) {
    switch (dtype) {
        case PO_U8: case PO_I8: return 1;
        case PO_U16: case PO_I16: return 2;
        case PO_U32: case PO_I32: case PO_F32: return 4;
        case PO_U64: case PO_I64: case PO_F64: return 8;
        case PO_VEC3U8: return 3;
        case PO_VEC3U16: return 6;
        case PO_VEC3I32: case PO_VEC3F32: return 12;
        case PO_VEC3F64: return 24;
        case PO_VEC4U8: return 4;
        case PO_BYTEARRAY: case PO_CUSTOM: return extra_size;
        default: return 0;
    }
}

/* point_layout.rs:100-126: align_of of the Rust type; nalgebra Vector3<T> has the alignment of T */
uint64_t po_dtype_min_alignment(uint32_t dtype, uint64_t extra_align) {
    switch (dtype) {
        case PO_U8: case PO_I8: case PO_VEC3U8: case PO_VEC4U8: case PO_BYTEARRAY: return 1;
        case PO_U16: case PO_I16: case PO_VEC3U16: return 2;
        case PO_U32: case PO_I32: case PO_F32: case PO_VEC3I32: case PO_VEC3F32: return 4;
        case PO_U64: case PO_I64: case PO_F64: case PO_VEC3F64: return 8;
        case PO_CUSTOM: return extra_align;
        default: return 1;
    }
}

/* math/arithmetic.rs:8 (Alignable::align_to) */
static uint64_t align_to(uint64_t v, uint64_t a) {
    if (a == 0) return v;
    uint64_t r = v % a;
    return r == 0 ? v : v + (a - r);
}

/* point_layout.rs:1011-1024 */
void po_layout_init(po_layout* l) {
    memset(l, 0, sizeof(*l));
    l->size = 0;
    l->align = 1;
}

int po_layout_index_by_name(const po_layout* l, const char* name) {
    for (uint32_t i = 0; i < l->n; ++i)
        if (strcmp(l->m[i].name, name) == 0) return (int)i;
    return -1;
}

static int member_dtype_equal(const po_member* a, uint32_t dtype, uint64_t extra_size) {
    if (a->dtype != dtype) return 0;
    if (dtype == PO_BYTEARRAY || dtype == PO_CUSTOM) return a->extra_size == extra_size;
    return 1;
}

/* point_layout.rs:951-956 (name + datatype) */
int po_layout_index_of(const po_layout* l, const char* name, uint32_t dtype) {
    for (uint32_t i = 0; i < l->n; ++i)
        if (strcmp(l->m[i].name, name) == 0 && l->m[i].dtype == dtype) return (int)i;
    return -1;
}

/* point_layout.rs:778-822 */
int po_layout_add_attribute(po_layout* l, const char* name, uint32_t dtype, uint64_t extra_size,
                            uint64_t extra_align, uint64_t packed_n) {
    if (l->n >= PO_MAX_ATTRS) return PO_ERR_INVALID;
    if (po_layout_index_by_name(l, name) >= 0) return PO_ERR_DUPLICATE_ATTR; /* :783-788 */
    uint64_t min_align = po_dtype_min_alignment(dtype, extra_align);
    uint64_t field_align = packed_n ? (packed_n < min_align ? packed_n : min_align) : min_align; /* :790-795 */
    uint64_t next = 0; /* :987-996 */
    if (l->n > 0) next = l->m[l->n - 1].offset + l->m[l->n - 1].size;
    uint64_t offset = align_to(next, field_align); /* :796-798 */
    uint64_t cur_max = l->align;
    uint64_t new_max = packed_n ? (packed_n < cur_max ? packed_n : cur_max) /* :806-808 */
                                : (cur_max > min_align ? cur_max : min_align); /* :802-805 */
    po_member* m = &l->m[l->n++];
    memset(m, 0, sizeof(*m));
    strncpy(m->name, name, PO_NAME_LEN - 1);
    m->dtype = dtype;
    m->extra_size = extra_size;
    m->extra_align = extra_align;
    m->offset = offset;
    m->size = po_dtype_size(dtype, extra_size);
    uint64_t end = offset + m->size; /* :814-821 */
    uint64_t unaligned = l->size > end ? l->size : end;
    l->size = align_to(unaligned, new_max);
    l->align = new_max;
    return PO_OK;
}

/* point_layout.rs:719-759 */
int po_layout_from_members_and_alignment(po_layout* l, const po_member* members, uint32_t n,
                                         uint64_t type_alignment) {
    if (n > PO_MAX_ATTRS) return PO_ERR_INVALID;
    po_layout_init(l);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j)
            if (strcmp(members[i].name, members[j].name) == 0) return PO_ERR_DUPLICATE_ATTR;
    { /* :732-743: sort ranges by start, adjacent ranges must not overlap */
        uint64_t st[PO_MAX_ATTRS], en[PO_MAX_ATTRS];
        for (uint32_t i = 0; i < n; ++i) { st[i] = members[i].offset; en[i] = members[i].offset + po_dtype_size(members[i].dtype, members[i].extra_size); }
        for (uint32_t i = 1; i < n; ++i) { /* insertion sort (stable, like sort_by) */
            uint64_t s0 = st[i], e0 = en[i]; uint32_t j = i;
            while (j > 0 && st[j - 1] > s0) { st[j] = st[j - 1]; en[j] = en[j - 1]; j--; }
            st[j] = s0; en[j] = e0;
        }
        for (uint32_t i = 1; i < n; ++i) if (en[i - 1] > st[i]) return PO_ERR_OVERLAP;
    }
    uint64_t unaligned = 0, max_off = 0;
    int have = 0;
    for (uint32_t i = 0; i < n; ++i) {
        po_member* m = &l->m[i];
        *m = members[i];
        m->name[PO_NAME_LEN - 1] = 0;
        m->size = po_dtype_size(m->dtype, m->extra_size);
        /* max_by offset: last maximal element wins; sizes only matter through offset+size */
        if (!have || m->offset >= max_off) {
            max_off = m->offset;
            unaligned = m->offset + m->size;
            have = 1;
        }
    }
    l->n = n;
    l->size = align_to(unaligned, type_alignment);
    l->align = type_alignment;
    return PO_OK;
}

/* derived PartialEq of PointLayout: attributes (name, dtype, offset, size) in order + memory layout */
int po_layout_equal(const po_layout* a, const po_layout* b) {
    if (a->n != b->n || a->size != b->size || a->align != b->align) return 0;
    for (uint32_t i = 0; i < a->n; ++i) {
        const po_member *x = &a->m[i], *y = &b->m[i];
        if (strcmp(x->name, y->name) != 0 || !member_dtype_equal(x, y->dtype, y->extra_size) ||
            x->offset != y->offset || x->size != y->size)
            return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * X: per-value casts with Rust `as` semantics (attribute_conversion.rs:184-343)
 * ---------------------------------------------------------------------------------------------- */

static int is_scalar(uint32_t d) { return d <= PO_F64; }
static int is_cast_vec3(uint32_t d) { return d == PO_VEC3U8 || d == PO_VEC3U16 || d == PO_VEC3F32 || d == PO_VEC3I32 || d == PO_VEC3F64; }
static uint32_t vec3_component(uint32_t d) {
    switch (d) {
        case PO_VEC3U8: return PO_U8;
        case PO_VEC3U16: return PO_U16;
        case PO_VEC3F32: return PO_F32;
        case PO_VEC3I32: return PO_I32;
        case PO_VEC3F64: return PO_F64;
        default: return PO_U8;
    }
}

/* attribute_conversion.rs:194-260: all ordered pairs of distinct scalars; all ordered pairs of distinct
 * cast-able Vec3 types. Same type -> no table entry (callers test equality first, :127,372). */
int po_has_conversion(uint32_t from, uint32_t to) {
    if (from == to) return 0;
    if (is_scalar(from) && is_scalar(to)) return 1;
    if (is_cast_vec3(from) && is_cast_vec3(to)) return 1;
    return 0;
}

/* Rust float -> int `as`: truncate toward zero, saturate, NaN -> 0 */
static int64_t f64_as_signed(double v, int bits) {
    if (v != v) return 0;
    double lim = ldexp(1.0, bits - 1);
    if (v >= lim) return bits == 64 ? INT64_MAX : (int64_t)((((uint64_t)1) << (bits - 1)) - 1);
    if (v <= -lim) return bits == 64 ? INT64_MIN : -(int64_t)(((uint64_t)1) << (bits - 1));
    return (int64_t)v;
}
static uint64_t f64_as_unsigned(double v, int bits) {
    if (v != v) return 0;
    if (v <= 0.0) return 0;
    double lim = ldexp(1.0, bits);
    if (v >= lim) return bits == 64 ? UINT64_MAX : ((((uint64_t)1) << bits) - 1);
    return (uint64_t)v;
}

typedef struct { int cls; int64_t i; uint64_t u; double f; float f32; } scalar_val; /* cls 0=signed 1=unsigned 2=f32 3=f64 */

static scalar_val load_scalar(uint32_t d, const uint8_t* p) {
    scalar_val v;
    memset(&v, 0, sizeof v);
    switch (d) {
        case PO_U8: v.cls = 1; v.u = p[0]; break;
        case PO_I8: v.cls = 0; v.i = (int8_t)p[0]; break;
        case PO_U16: { uint16_t x; memcpy(&x, p, 2); v.cls = 1; v.u = x; break; }
        case PO_I16: { int16_t x; memcpy(&x, p, 2); v.cls = 0; v.i = x; break; }
        case PO_U32: { uint32_t x; memcpy(&x, p, 4); v.cls = 1; v.u = x; break; }
        case PO_I32: { int32_t x; memcpy(&x, p, 4); v.cls = 0; v.i = x; break; }
        case PO_U64: { uint64_t x; memcpy(&x, p, 8); v.cls = 1; v.u = x; break; }
        case PO_I64: { int64_t x; memcpy(&x, p, 8); v.cls = 0; v.i = x; break; }
        case PO_F32: { float x; memcpy(&x, p, 4); v.cls = 2; v.f32 = x; v.f = (double)x; break; }
        case PO_F64: { double x; memcpy(&x, p, 8); v.cls = 3; v.f = x; break; }
        default: break;
    }
    return v;
}

/* attribute_conversion.rs:310-321 (`from_value.as_()` == Rust `as`) */
static void cast_scalar(uint32_t from, uint32_t to, const uint8_t* src, uint8_t* dst) {
    scalar_val v = load_scalar(from, src);
    int is_float_src = v.cls >= 2;
    switch (to) {
        case PO_U8: case PO_U16: case PO_U32: case PO_U64: {
            int bits = to == PO_U8 ? 8 : to == PO_U16 ? 16 : to == PO_U32 ? 32 : 64;
            uint64_t r;
            if (is_float_src) r = f64_as_unsigned(v.f, bits); /* f32 -> f64 is exact, so this equals f32 `as` */
            else r = v.cls == 0 ? (uint64_t)v.i : v.u;         /* wrap / sign-extend */
            if (to == PO_U8) { uint8_t x = (uint8_t)r; memcpy(dst, &x, 1); }
            else if (to == PO_U16) { uint16_t x = (uint16_t)r; memcpy(dst, &x, 2); }
            else if (to == PO_U32) { uint32_t x = (uint32_t)r; memcpy(dst, &x, 4); }
            else memcpy(dst, &r, 8);
            break;
        }
        case PO_I8: case PO_I16: case PO_I32: case PO_I64: {
            int bits = to == PO_I8 ? 8 : to == PO_I16 ? 16 : to == PO_I32 ? 32 : 64;
            uint64_t r;
            if (is_float_src) r = (uint64_t)f64_as_signed(v.f, bits);
            else r = v.cls == 0 ? (uint64_t)v.i : v.u;
            if (to == PO_I8) { uint8_t x = (uint8_t)r; memcpy(dst, &x, 1); }
            else if (to == PO_I16) { uint16_t x = (uint16_t)r; memcpy(dst, &x, 2); }
            else if (to == PO_I32) { uint32_t x = (uint32_t)r; memcpy(dst, &x, 4); }
            else memcpy(dst, &r, 8);
            break;
        }
        case PO_F32: {
            float r;
            if (v.cls == 0) r = (float)v.i;       /* RNE */
            else if (v.cls == 1) r = (float)v.u;  /* RNE */
            else if (v.cls == 2) r = v.f32;
            else r = (float)v.f;                  /* RNE, overflow -> inf */
            memcpy(dst, &r, 4);
            break;
        }
        case PO_F64: {
            double r;
            if (v.cls == 0) r = (double)v.i;
            else if (v.cls == 1) r = (double)v.u;
            else r = v.f;
            memcpy(dst, &r, 8);
            break;
        }
        default: break;
    }
}

int po_convert_value(uint32_t from, uint32_t to, const uint8_t* src, uint8_t* dst) {
    if (!po_has_conversion(from, to)) return PO_ERR_NO_CONVERSION;
    if (is_scalar(from)) {
        cast_scalar(from, to, src, dst);
    } else { /* attribute_conversion.rs:332-343 */
        uint32_t cf = vec3_component(from), ct = vec3_component(to);
        uint64_t sf = po_dtype_size(cf, 0), st = po_dtype_size(ct, 0);
        for (int c = 0; c < 3; ++c) {
            if (cf == ct) memcpy(dst + c * st, src + c * sf, sf);
            else cast_scalar(cf, ct, src + c * sf, dst + c * st);
        }
    }
    return PO_OK;
}

/* function-pointer style converter, as the reference stores it (attribute_conversion.rs:112) */
typedef void (*conv_fn)(uint32_t from, uint32_t to, const uint8_t* src, uint8_t* dst);
static void conv_dispatch(uint32_t from, uint32_t to, const uint8_t* src, uint8_t* dst) {
    (void)po_convert_value(from, to, src, dst);
}

/* ------------------------------------------------------------------------------------------------
 * T: enumerated transforms = the closures used in-tree
 * ---------------------------------------------------------------------------------------------- */

/* which dtypes a transform kind is defined on */
static int transform_supports(uint32_t kind, uint32_t dtype) {
    switch (kind) {
        case PO_T_SCALE_OFFSET: case PO_T_ADD:
            return dtype == PO_VEC3F64 || dtype == PO_VEC3F32 || dtype == PO_F64 || dtype == PO_F32;
        case PO_T_INV_SCALE_OFFSET:
            return dtype == PO_VEC3F64 || dtype == PO_F64;
        case PO_T_SHIFT_MASK:
            return dtype == PO_U8 || dtype == PO_U16 || dtype == PO_U32 || dtype == PO_U64;
        default: return 0;
    }
}

static void apply_transform(const po_transform* t, uint32_t dtype, uint8_t* mem) {
    switch (t->kind) {
        case PO_T_SCALE_OFFSET:
            if (dtype == PO_VEC3F64 || dtype == PO_F64) { /* raw_readers.rs:42-48: (pos*scale)+offset, two roundings */
                int n = dtype == PO_VEC3F64 ? 3 : 1;
                for (int c = 0; c < n; ++c) { double v; memcpy(&v, mem + 8 * c, 8); double m = v * t->s[c]; v = m + t->o[c]; memcpy(mem + 8 * c, &v, 8); }
            } else { /* raw_readers.rs:49-55: ((pos as f64*scale)+offset) as f32 */
                int n = dtype == PO_VEC3F32 ? 3 : 1;
                for (int c = 0; c < n; ++c) { float v; memcpy(&v, mem + 4 * c, 4); double m = (double)v * t->s[c]; double r = m + t->o[c]; v = (float)r; memcpy(mem + 4 * c, &v, 4); }
            }
            break;
        case PO_T_INV_SCALE_OFFSET: { /* write_helpers.rs:15-17: (p - offset) / scale */
            int n = dtype == PO_VEC3F64 ? 3 : 1;
            for (int c = 0; c < n; ++c) { double v; memcpy(&v, mem + 8 * c, 8); double d = v - t->o[c]; v = d / t->s[c]; memcpy(mem + 8 * c, &v, 8); }
            break;
        }
        case PO_T_ADD:
            if (dtype == PO_VEC3F64 || dtype == PO_F64) { /* pnts_reader.rs:275-277 ; buffer_conversion.rs:780-782 */
                int n = dtype == PO_VEC3F64 ? 3 : 1;
                for (int c = 0; c < n; ++c) { double v; memcpy(&v, mem + 8 * c, 8); v = v + t->o[c]; memcpy(mem + 8 * c, &v, 8); }
            } else { /* pnts_reader.rs:265-273 */
                int n = dtype == PO_VEC3F32 ? 3 : 1;
                for (int c = 0; c < n; ++c) { float v; memcpy(&v, mem + 4 * c, 4); double r = (double)v + t->o[c]; v = (float)r; memcpy(mem + 4 * c, &v, 4); }
            }
            break;
        case PO_T_SHIFT_MASK: { /* raw_readers.rs:61-164 */
            switch (dtype) {
                case PO_U8: { uint8_t v = mem[0]; v = (uint8_t)((v >> t->shift) & t->mask); mem[0] = v; break; }
                case PO_U16: { uint16_t v; memcpy(&v, mem, 2); v = (uint16_t)((v >> t->shift) & t->mask); memcpy(mem, &v, 2); break; }
                case PO_U32: { uint32_t v; memcpy(&v, mem, 4); v = (uint32_t)((v >> t->shift) & t->mask); memcpy(mem, &v, 4); break; }
                case PO_U64: { uint64_t v; memcpy(&v, mem, 8); v = (v >> t->shift) & t->mask; memcpy(mem, &v, 8); break; }
                default: break;
            }
            break;
        }
        default: break;
    }
}

/* ------------------------------------------------------------------------------------------------
 * C: BufferLayoutConverter (layout/conversion/buffer_conversion.rs)
 * ---------------------------------------------------------------------------------------------- */

/* buffer_conversion.rs:368-396 */
static int make_default_mapping(const po_layout* from, int si, const po_layout* to, int ti, po_mapping* out) {
    memset(out, 0, sizeof(*out));
    out->source_idx = si;
    out->target_idx = ti;
    const po_member *fm = &from->m[si], *tm = &to->m[ti];
    if (member_dtype_equal(fm, tm->dtype, tm->extra_size)) {
        out->has_converter = 0;
    } else {
        if (!po_has_conversion(fm->dtype, tm->dtype)) return PO_ERR_NO_CONVERSION; /* :383-388 panic */
        out->has_converter = 1;
    }
    return PO_OK;
}

/* buffer_conversion.rs:112-143 */
int po_converter_for_layouts(po_converter* cv, const po_layout* from, const po_layout* to, int with_default) {
    memset(cv, 0, sizeof(*cv));
    cv->from_layout = *from;
    cv->to_layout = *to;
    for (uint32_t t = 0; t < to->n; ++t) {
        int s = po_layout_index_by_name(from, to->m[t].name);
        if (s < 0) {
            if (with_default) continue;     /* :132-136 filter_map */
            return PO_ERR_ATTR_NOT_FOUND;   /* :114 expect */
        }
        int rc = make_default_mapping(&cv->from_layout, s, &cv->to_layout, (int)t, &cv->mappings[cv->n_mappings]);
        if (rc) return rc;
        cv->n_mappings++;
    }
    return PO_OK;
}

static int find_mapping_for_target(po_converter* cv, int ti) {
    for (uint32_t i = 0; i < cv->n_mappings; ++i)
        if (cv->mappings[i].target_idx == ti) return (int)i;
    return -1;
}

/* buffer_conversion.rs:156-183 */
int po_converter_set_custom_mapping(po_converter* cv, const char* from_name, uint32_t from_dtype,
                                    const char* to_name, uint32_t to_dtype) {
    int s = po_layout_index_of(&cv->from_layout, from_name, from_dtype);
    if (s < 0) return PO_ERR_ATTR_NOT_FOUND;
    int t = po_layout_index_of(&cv->to_layout, to_name, to_dtype);
    if (t < 0) return PO_ERR_ATTR_NOT_FOUND;
    po_mapping m;
    int rc = make_default_mapping(&cv->from_layout, s, &cv->to_layout, t, &m);
    if (rc) return rc;
    int prev = find_mapping_for_target(cv, t);
    if (prev >= 0) cv->mappings[prev] = m;
    else cv->mappings[cv->n_mappings++] = m;
    return PO_OK;
}

/* buffer_conversion.rs:194-234, 404-416 */
int po_converter_set_custom_mapping_with_transformation(po_converter* cv, const char* from_name,
                                                        uint32_t from_dtype, const char* to_name,
                                                        uint32_t to_dtype, uint32_t transform_dtype,
                                                        const po_transform* tr, int apply_to_source) {
    int s = po_layout_index_of(&cv->from_layout, from_name, from_dtype);
    if (s < 0) return PO_ERR_ATTR_NOT_FOUND;
    int t = po_layout_index_of(&cv->to_layout, to_name, to_dtype);
    if (t < 0) return PO_ERR_ATTR_NOT_FOUND;
    /* :209-213 assert_eq!(T::data_type(), ...) */
    if (apply_to_source ? transform_dtype != cv->from_layout.m[s].dtype : transform_dtype != cv->to_layout.m[t].dtype)
        return PO_ERR_TRANSFORM_DTYPE;
    if (!transform_supports(tr->kind, transform_dtype)) return PO_ERR_UNSUPPORTED;
    po_mapping m;
    int rc = make_default_mapping(&cv->from_layout, s, &cv->to_layout, t, &m);
    if (rc) return rc;
    m.has_transform = 1;
    m.apply_to_source = apply_to_source ? 1 : 0;
    m.t = *tr;
    int prev = find_mapping_for_target(cv, t);
    if (prev >= 0) cv->mappings[prev] = m;
    else cv->mappings[cv->n_mappings++] = m;
    return PO_OK;
}

/* raw_attribute_view.rs:19-71: (base, offset, stride, size) */
typedef struct { uint8_t* base; uint64_t offset, stride, size; } raw_view;

static raw_view view_attr(const po_buffer* b, int idx) {
    raw_view v;
    const po_member* m = &b->layout->m[idx];
    if (b->columnar) { v.base = b->columns[idx]; v.offset = 0; v.stride = m->size; v.size = m->size; }
    else { v.base = b->aos; v.offset = m->offset; v.stride = b->layout->size; v.size = m->size; }
    return v;
}
static inline uint8_t* view_at(const raw_view* v, uint64_t i) { return v->base + v->offset + v->stride * i; }

/* The four loops of buffer_conversion.rs:418-662 share one per-element rule; the loop nest
 * (mapping outer, point inner) and the scratch-copy behaviour are kept as in the reference. */
static void convert_mappings(const po_converter* cv, const po_buffer* src, uint64_t sb, po_buffer* dst,
                             uint64_t db, uint64_t count, uint32_t map_begin, uint32_t map_end) {
    uint8_t scratch[64];
    uint8_t* big = NULL;
    volatile conv_fn converter = conv_dispatch; /* indirect call per value, like the reference's fn pointer */
    for (uint32_t mi = map_begin; mi < map_end; ++mi) {
        const po_mapping* mp = &cv->mappings[mi];
        const po_member* sm = &cv->from_layout.m[mp->source_idx];
        const po_member* tm = &cv->to_layout.m[mp->target_idx];
        raw_view sv = view_attr(src, mp->source_idx);
        raw_view tv = view_attr(dst, mp->target_idx);
        uint8_t* buf = scratch;
        if (sm->size > sizeof scratch) { big = (uint8_t*)realloc(big, sm->size); buf = big; }
        for (uint64_t i = 0; i < count; ++i) {
            const uint8_t* s = view_at(&sv, sb + i);
            uint8_t* t = view_at(&tv, db + i);
            if (mp->has_converter) {
                if (mp->has_transform) {
                    if (mp->apply_to_source) { /* :573-579 */
                        memcpy(buf, s, sm->size);
                        apply_transform(&mp->t, sm->dtype, buf);
                        converter(sm->dtype, tm->dtype, buf, t);
                    } else { /* :581-584 */
                        converter(sm->dtype, tm->dtype, s, t);
                        apply_transform(&mp->t, tm->dtype, t);
                    }
                } else {
                    converter(sm->dtype, tm->dtype, s, t); /* :590-592 */
                }
            } else if (mp->has_transform) { /* :594-598 (same dtype on both sides) */
                memcpy(buf, s, sm->size);
                apply_transform(&mp->t, sm->dtype, buf);
                memcpy(t, buf, sm->size);
            } else {
                memcpy(t, s, sm->size); /* :600 */
            }
        }
    }
    free(big);
}

/* buffer_conversion.rs:292-359 */
int po_convert_into_range(const po_converter* cv, const po_buffer* src, uint64_t sb, uint64_t se,
                          po_buffer* dst, uint64_t db, uint64_t de) {
    if (!po_layout_equal(src->layout, &cv->from_layout)) return PO_ERR_LAYOUT_MISMATCH; /* :302 */
    if (!po_layout_equal(dst->layout, &cv->to_layout)) return PO_ERR_LAYOUT_MISMATCH;   /* :303 */
    if (se < sb || de < db || se - sb != de - db) return PO_ERR_RANGE;                  /* :304 */
    if (se > src->len) return PO_ERR_RANGE;                                             /* :305 */
    if (de > dst->len) return PO_ERR_RANGE;                                             /* :306 */
    convert_mappings(cv, src, sb, dst, db, se - sb, 0, cv->n_mappings);
    return PO_OK;
}

typedef struct { const po_converter* cv; const po_buffer* src; po_buffer* dst; uint64_t sb, db, count; } mt_job;
static void* mt_worker(void* p) {
    mt_job* j = (mt_job*)p;
    convert_mappings(j->cv, j->src, j->sb, j->dst, j->db, j->count, 0, j->cv->n_mappings);
    return NULL;
}

int po_convert_into_range_mt(const po_converter* cv, const po_buffer* src, uint64_t sb, uint64_t se,
                             po_buffer* dst, uint64_t db, uint64_t de, int n_threads) {
    if (!po_layout_equal(src->layout, &cv->from_layout)) return PO_ERR_LAYOUT_MISMATCH;
    if (!po_layout_equal(dst->layout, &cv->to_layout)) return PO_ERR_LAYOUT_MISMATCH;
    if (se < sb || de < db || se - sb != de - db || se > src->len || de > dst->len) return PO_ERR_RANGE;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    uint64_t count = se - sb;
    pthread_t th[256];
    mt_job jobs[256];
    uint64_t per = (count + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    int started = 0;
    for (int i = 0; i < n_threads; ++i) {
        uint64_t b = per * (uint64_t)i;
        if (b >= count) break;
        uint64_t e = b + per > count ? count : b + per;
        jobs[i] = (mt_job){cv, src, dst, sb + b, db + b, e - b};
        pthread_create(&th[i], NULL, mt_worker, &jobs[i]);
        started++;
    }
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return PO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * T: LAS layouts and the default LAS converter
 * ---------------------------------------------------------------------------------------------- */

typedef struct { int extended, gps, color, nir, waveform; } las_format;
static int las_format_of(int n, las_format* f) { /* las::point::Format::new(n) */
    if (n < 0 || n > 10) return PO_ERR_INVALID;
    f->extended = n >= 6;
    f->gps = (n == 1 || n == 3 || n == 4 || n == 5 || n >= 6);
    f->color = (n == 2 || n == 3 || n == 5 || n == 7 || n == 8 || n == 10);
    f->nir = (n == 8 || n == 10);
    f->waveform = (n == 4 || n == 5 || n == 9 || n == 10);
    return PO_OK;
}

/* las_layout.rs:70-108 (exact_binary_representation == true) */
int po_las_raw_layout(int format, po_layout* l) {
    las_format f;
    int rc = las_format_of(format, &f);
    if (rc) return rc;
    po_layout_init(l);
    po_layout_add_attribute(l, "LASLocalPosition", PO_VEC3I32, 0, 0, 1);
    po_layout_add_attribute(l, "Intensity", PO_U16, 0, 0, 1);
    if (f.extended) po_layout_add_attribute(l, "LASExtendedFlags", PO_U16, 0, 0, 1);
    else po_layout_add_attribute(l, "LASBasicFlags", PO_U8, 0, 0, 1);
    po_layout_add_attribute(l, "Classification", PO_U8, 0, 0, 1);
    if (f.extended) {
        po_layout_add_attribute(l, "UserData", PO_U8, 0, 0, 1);
        po_layout_add_attribute(l, "ScanAngle", PO_I16, 0, 0, 1);
    } else {
        po_layout_add_attribute(l, "ScanAngleRank", PO_I8, 0, 0, 1);
        po_layout_add_attribute(l, "UserData", PO_U8, 0, 0, 1);
    }
    po_layout_add_attribute(l, "PointSourceID", PO_U16, 0, 0, 1);
    if (f.gps) po_layout_add_attribute(l, "GpsTime", PO_F64, 0, 0, 1);
    if (f.color) po_layout_add_attribute(l, "ColorRGB", PO_VEC3U16, 0, 0, 1);
    if (f.nir) po_layout_add_attribute(l, "NIR", PO_U16, 0, 0, 1);
    if (f.waveform) {
        po_layout_add_attribute(l, "WavePacketDescriptorIndex", PO_U8, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformDataOffset", PO_U64, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformPacketSize", PO_U32, 0, 0, 1);
        po_layout_add_attribute(l, "ReturnPointWaveformLocation", PO_F32, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformParameters", PO_VEC3F32, 0, 0, 1);
    }
    return PO_OK;
}

/* las_types.rs: #[repr(C, packed)] structs LasPointFormat0..10 (derive(PointType) -> packed(1) offsets) */
int po_las_default_layout(int format, po_layout* l) {
    las_format f;
    int rc = las_format_of(format, &f);
    if (rc) return rc;
    po_layout_init(l);
    po_layout_add_attribute(l, "Position3D", PO_VEC3F64, 0, 0, 1);
    po_layout_add_attribute(l, "Intensity", PO_U16, 0, 0, 1);
    po_layout_add_attribute(l, "ReturnNumber", PO_U8, 0, 0, 1);
    po_layout_add_attribute(l, "NumberOfReturns", PO_U8, 0, 0, 1);
    if (f.extended) {
        po_layout_add_attribute(l, "ClassificationFlags", PO_U8, 0, 0, 1);
        po_layout_add_attribute(l, "ScannerChannel", PO_U8, 0, 0, 1);
    }
    po_layout_add_attribute(l, "ScanDirectionFlag", PO_U8, 0, 0, 1);
    po_layout_add_attribute(l, "EdgeOfFlightLine", PO_U8, 0, 0, 1);
    po_layout_add_attribute(l, "Classification", PO_U8, 0, 0, 1);
    if (f.extended) {
        po_layout_add_attribute(l, "UserData", PO_U8, 0, 0, 1);
        po_layout_add_attribute(l, "ScanAngle", PO_I16, 0, 0, 1);
    } else {
        po_layout_add_attribute(l, "ScanAngleRank", PO_I8, 0, 0, 1);
        po_layout_add_attribute(l, "UserData", PO_U8, 0, 0, 1);
    }
    po_layout_add_attribute(l, "PointSourceID", PO_U16, 0, 0, 1);
    if (f.gps) po_layout_add_attribute(l, "GpsTime", PO_F64, 0, 0, 1);
    if (f.color) po_layout_add_attribute(l, "ColorRGB", PO_VEC3U16, 0, 0, 1);
    if (f.nir) po_layout_add_attribute(l, "NIR", PO_U16, 0, 0, 1);
    if (f.waveform) {
        po_layout_add_attribute(l, "WavePacketDescriptorIndex", PO_U8, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformDataOffset", PO_U64, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformPacketSize", PO_U32, 0, 0, 1);
        po_layout_add_attribute(l, "ReturnPointWaveformLocation", PO_F32, 0, 0, 1);
        po_layout_add_attribute(l, "WaveformParameters", PO_VEC3F32, 0, 0, 1);
    }
    return PO_OK;
}

static int add_bitfield(po_converter* cv, const char* flags_name, uint32_t flags_dtype, const po_layout* target,
                        const char* target_name, uint32_t shift, uint64_t mask) {
    int ti = po_layout_index_by_name(target, target_name);
    if (ti < 0) return PO_OK;
    po_transform t;
    memset(&t, 0, sizeof t);
    t.kind = PO_T_SHIFT_MASK;
    t.shift = shift;
    t.mask = mask;
    return po_converter_set_custom_mapping_with_transformation(cv, flags_name, flags_dtype, target_name,
                                                               target->m[ti].dtype, flags_dtype, &t, 1);
}

/* raw_readers.rs:31-167 */
int po_las_default_converter(po_converter* cv, const po_layout* raw, const po_layout* target,
                             const double scale[3], const double offset[3]) {
    int rc = po_converter_for_layouts(cv, raw, target, 1);
    if (rc) return rc;
    int pi = po_layout_index_by_name(target, "Position3D");
    if (pi >= 0) {
        uint32_t d = target->m[pi].dtype;
        if (d != PO_VEC3F64 && d != PO_VEC3F32) return PO_ERR_UNSUPPORTED; /* :56 bail! */
        po_transform t;
        memset(&t, 0, sizeof t);
        t.kind = PO_T_SCALE_OFFSET;
        for (int c = 0; c < 3; ++c) { t.s[c] = scale[c]; t.o[c] = offset[c]; }
        rc = po_converter_set_custom_mapping_with_transformation(cv, "LASLocalPosition", PO_VEC3I32, "Position3D", d, d, &t, 0);
        if (rc) return rc;
    }
    if (po_layout_index_of(raw, "LASBasicFlags", PO_U8) >= 0) { /* :61-103 */
        if ((rc = add_bitfield(cv, "LASBasicFlags", PO_U8, target, "ReturnNumber", 0, 0x7))) return rc;
        if ((rc = add_bitfield(cv, "LASBasicFlags", PO_U8, target, "NumberOfReturns", 3, 0x7))) return rc;
        if ((rc = add_bitfield(cv, "LASBasicFlags", PO_U8, target, "ScanDirectionFlag", 6, 0x1))) return rc;
        if ((rc = add_bitfield(cv, "LASBasicFlags", PO_U8, target, "EdgeOfFlightLine", 7, 0x1))) return rc;
    } else { /* :104-164 */
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "ReturnNumber", 0, 0xF))) return rc;
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "NumberOfReturns", 4, 0xF))) return rc;
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "ClassificationFlags", 8, 0xF))) return rc;
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "ScannerChannel", 12, 0x3))) return rc;
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "ScanDirectionFlag", 14, 0x1))) return rc;
        if ((rc = add_bitfield(cv, "LASExtendedFlags", PO_U16, target, "EdgeOfFlightLine", 15, 0x1))) return rc;
    }
    return PO_OK;
}

/* write_helpers.rs:10-23: (((p - off) / scale) as i64).try_into::<i32>().expect(..) */
int po_las_write_position(const double world[3], const double scale[3], const double offset[3], int32_t out[3]) {
    int panic = 0;
    for (int c = 0; c < 3; ++c) {
        double d = world[c] - offset[c];
        double q = d / scale[c];
        int64_t l = f64_as_signed(q, 64);
        if (l > INT32_MAX || l < INT32_MIN) { panic = 1; out[c] = l > 0 ? INT32_MAX : INT32_MIN; }
        else out[c] = (int32_t)l;
    }
    return panic;
}

/* RawLASWriter::write_points_default_layout, pasture-io/src/las/raw_writers.rs:203-362. `src` must have the default
 * layout of `format` (LasPointFormatN). out: n raw records. counts16[r] = points with return number r (r = 1..15; the
 * reference keeps 1..=5 or 1..=15 depending on the header, :221-229). Returns the number of points for which
 * write_position_as_las_position would panic (write_helpers.rs:15-17); bounds as update_bounds_in_las_header (:28-47)
 * starting from (+MAX, -MAX). */
int po_las_write_points(const po_buffer* src, int format, const double scale[3], const double offset[3], uint8_t* out,
                        uint64_t counts16[16], double bmin[3], double bmax[3], uint64_t* panics) {
    po_layout def, raw;
    int rc = po_las_default_layout(format, &def);
    if (rc) return rc;
    if (!po_layout_equal(src->layout, &def)) return PO_ERR_LAYOUT_MISMATCH;
    po_las_raw_layout(format, &raw);
    las_format f;
    las_format_of(format, &f);
    for (int b = 0; b < 16; ++b) counts16[b] = 0;
    for (int c = 0; c < 3; ++c) { bmin[c] = 1.7976931348623157e308; bmax[c] = -1.7976931348623157e308; }
    *panics = 0;
    for (uint64_t i = 0; i < src->len; ++i) {
        uint8_t rec[128];
        /* get_point_range: gather the point in default-layout order (:237-240) */
        for (uint32_t a = 0; a < def.n; ++a) {
            raw_view v = view_attr(src, (int)a);
            memcpy(rec + def.m[a].offset, view_at(&v, i), def.m[a].size);
        }
        const uint8_t* r = rec;
        uint8_t* w = out + i * raw.size;
        double pos[3];
        memcpy(pos, r, 24); r += 24;
        int32_t local[3];
        if (po_las_write_position(pos, scale, offset, local)) (*panics)++;
        memcpy(w, local, 12); w += 12;
        for (int c = 0; c < 3; ++c) { if (pos[c] < bmin[c]) bmin[c] = pos[c]; if (pos[c] > bmax[c]) bmax[c] = pos[c]; }
        memcpy(w, r, 2); w += 2; r += 2; /* intensity */
        if (f.extended) {
            uint8_t rn = r[0], nr = r[1], cf = r[2], sc = r[3], sd = r[4], ef = r[5];
            r += 6;
            if (rn >= 1 && rn <= 15) counts16[rn]++;
            w[0] = (uint8_t)((rn & 0xF) | ((nr & 0xF) << 4));                                   /* write_helpers.rs:39-47 */
            w[1] = (uint8_t)((cf & 0xF) | ((sc & 0x3) << 4) | ((sd & 0x1) << 6) | ((ef & 0x1) << 7));
            w += 2;
        } else {
            uint8_t rn = r[0], nr = r[1], sd = r[2], ef = r[3];
            r += 4;
            if (rn >= 1 && rn <= 15) counts16[rn]++;
            w[0] = (uint8_t)((rn & 0x7) | ((nr & 0x7) << 3) | ((sd & 0x1) << 6) | ((ef & 0x1) << 7)); /* :32-37 */
            w += 1;
        }
        *w++ = *r++; /* classification */
        if (f.extended) { *w++ = r[0]; memcpy(w, r + 1, 2); w += 2; r += 3; } /* user data, scan angle i16 */
        else { *w++ = r[0]; *w++ = r[1]; r += 2; }                              /* scan angle rank, user data */
        memcpy(w, r, 2); w += 2; r += 2; /* point source id */
        if (f.gps) { memcpy(w, r, 8); w += 8; r += 8; }
        if (f.color) { memcpy(w, r, 6); w += 6; r += 6; }
        if (f.nir) { memcpy(w, r, 2); w += 2; r += 2; }
        if (f.waveform) { memcpy(w, r, 29); w += 29; r += 29; }
    }
    return PO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * A / M: bounds and min-max
 * ---------------------------------------------------------------------------------------------- */

/* bounds.rs:11-85 */
int po_calculate_bounds(const po_buffer* buf, double out_min[3], double out_max[3]) {
    if (buf->len == 0) return 0;
    int pi = po_layout_index_by_name(buf->layout, "Position3D");
    if (pi < 0) return 0;
    uint32_t d = buf->layout->m[pi].dtype;
    if (d != PO_VEC3F64 && !po_has_conversion(d, PO_VEC3F64)) return 0; /* get_generic_converter would panic; treat as None */
    double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    raw_view v = view_attr(buf, pi);
    for (uint64_t i = 0; i < buf->len; ++i) {
        double p[3];
        if (d == PO_VEC3F64) memcpy(p, view_at(&v, i), 24);
        else po_convert_value(d, PO_VEC3F64, view_at(&v, i), (uint8_t*)p); /* :62-65 converting view */
        for (int c = 0; c < 3; ++c) {
            if (p[c] < mn[c]) mn[c] = p[c]; /* :34-51 strict compares: NaN never updates */
            if (p[c] > mx[c]) mx[c] = p[c];
        }
    }
    /* AABB::from_min_max panics if min > max (math/bounds.rs:21-26): only possible when every value is NaN */
    if (mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]) return PO_ERR_INVALID;
    memcpy(out_min, mn, 24);
    memcpy(out_max, mx, 24);
    return 1;
}

/* math/minmax.rs:36-112: ints cmp::min/max; floats `if a < b {a} else {b}` with self = new value */
static void minmax_scalar(uint32_t d, const uint8_t* val, uint8_t* mn, uint8_t* mx) {
    scalar_val v = load_scalar(d, val), a = load_scalar(d, mn), b = load_scalar(d, mx);
    uint64_t sz = po_dtype_size(d, 0);
    switch (v.cls) {
        case 0: if (v.i < a.i) memcpy(mn, val, sz); if (v.i > b.i) memcpy(mx, val, sz); break;
        case 1: if (v.u < a.u) memcpy(mn, val, sz); if (v.u > b.u) memcpy(mx, val, sz); break;
        case 2: /* val.infimum(&old): if val < old {val} else {old} */
            if (v.f32 < a.f32) memcpy(mn, val, sz);
            if (v.f32 > b.f32) memcpy(mx, val, sz);
            break;
        default:
            if (v.f < a.f) memcpy(mn, val, sz);
            if (v.f > b.f) memcpy(mx, val, sz);
            break;
    }
}

/* minmax.rs:13-51 */
int po_minmax_attribute(const po_buffer* buf, const char* name, uint32_t attr_dtype, uint32_t view_dtype,
                        uint8_t* out_min, uint8_t* out_max) {
    if (po_layout_index_by_name(buf->layout, name) < 0) return PO_ERR_ATTR_NOT_FOUND; /* :17-26 panic */
    if (view_dtype != attr_dtype) return PO_ERR_TRANSFORM_DTYPE; /* buffer_views.rs:549 assert_eq -> panic */
    int idx = po_layout_index_of(buf->layout, name, attr_dtype);
    if (idx < 0) return PO_ERR_ATTR_NOT_FOUND; /* view_attribute::<T> panics */
    if (!(is_scalar(attr_dtype) || is_cast_vec3(attr_dtype))) return PO_ERR_UNSUPPORTED; /* no MinMax impl */
    if (buf->len == 0) return 0;
    raw_view v = view_attr(buf, idx);
    uint64_t sz = v.size;
    memcpy(out_min, view_at(&v, 0), sz);
    memcpy(out_max, view_at(&v, 0), sz);
    uint32_t comp = is_scalar(attr_dtype) ? attr_dtype : vec3_component(attr_dtype);
    uint64_t cs = po_dtype_size(comp, 0);
    int nc = is_scalar(attr_dtype) ? 1 : 3;
    for (uint64_t i = 1; i < buf->len; ++i) {
        const uint8_t* p = view_at(&v, i);
        for (int c = 0; c < nc; ++c) minmax_scalar(comp, p + c * cs, out_min + c * cs, out_max + c * cs);
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Z: bit manipulation (math/bitmanip.rs)
 * ---------------------------------------------------------------------------------------------- */

/* bitmanip.rs:2-10 */
uint64_t po_expand_bits_by_3(uint64_t val) {
    val &= 0x1FFFFFull;
    val = (val | (val << 32)) & 0x00FF00000000FFFFull;
    val = (val | (val << 16)) & 0x00FF0000FF0000FFull;
    val = (val | (val << 8)) & 0xF00F00F00F00F00Full;
    val = (val | (val << 4)) & 0x30C30C30C30C30C3ull;
    val = (val | (val << 2)) & 0x1249249249249249ull;
    return val;
}

/* bitmanip.rs:32-41 */
uint64_t po_reverse_bits(uint64_t v) {
    uint64_t r = 0;
    for (int i = 0; i < 64; ++i) r |= ((v >> i) & 1ull) << (63 - i);
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * V: voxel grid filter (pasture-algorithms/src/voxel_grid.rs)
 * ---------------------------------------------------------------------------------------------- */

/* voxel_grid.rs:54-79: running sum, `while cur < max { cur += leaf; push(cur) }` */
uint64_t po_create_markers(double bmin, double bmax, double leaf, double* out, uint64_t cap) {
    uint64_t n = 0;
    double cur = bmin;
    while (cur < bmax) {
        cur += leaf;
        if (n < cap) out[n] = cur;
        n++;
        if (!(leaf > 0.0) && n > cap) break; /* guard: non-positive leaf would loop forever in the reference */
    }
    return n;
}

static uint64_t leaf_axis_linear(double p, const double* m, uint64_t n) {
    uint64_t i = 0;
    while (n != 0 && m[i] < p) i++; /* :31-39; the reference would index out of bounds if p > last marker */
    if (i > 0 && p - m[i - 1] < m[i] - p) i--; /* :41-49 */
    return i;
}

/* voxel_grid.rs:22-51 */
void po_find_leaf(const double p[3], const double* mx, uint64_t nx, const double* my, uint64_t ny,
                  const double* mz, uint64_t nz, uint64_t out[3]) {
    out[0] = leaf_axis_linear(p[0], mx, nx);
    out[1] = leaf_axis_linear(p[1], my, ny);
    out[2] = leaf_axis_linear(p[2], mz, nz);
}

static uint64_t leaf_axis_bsearch(double p, const double* m, uint64_t n) {
    /* first index with !(m[i] < p); markers are strictly increasing for leaf > 0 */
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (m[mid] < p) lo = mid + 1; else hi = mid;
    }
    uint64_t i = lo;
    if (n == 0) return 0;
    if (i > 0 && p - m[i - 1] < m[i] - p) i--;
    return i;
}

void po_find_leaf_bsearch(const double p[3], const double* mx, uint64_t nx, const double* my, uint64_t ny,
                          const double* mz, uint64_t nz, uint64_t out[3]) {
    out[0] = leaf_axis_bsearch(p[0], mx, nx);
    out[1] = leaf_axis_bsearch(p[1], my, ny);
    out[2] = leaf_axis_bsearch(p[2], mz, nz);
}

typedef struct { uint64_t pos[3]; uint64_t* pts; uint64_t n, cap; } voxel_t;

static int cmp_pos(const uint64_t a[3], const uint64_t b[3]) {
    for (int c = 0; c < 3; ++c) { if (a[c] < b[c]) return -1; if (a[c] > b[c]) return 1; }
    return 0;
}

static void voxel_push(voxel_t* v, uint64_t i) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 4; v->pts = (uint64_t*)realloc(v->pts, v->cap * sizeof(uint64_t)); }
    v->pts[v->n++] = i;
}

static double attr_as_f64(const po_buffer* b, int idx, uint64_t p, int comp) {
    raw_view v = view_attr(b, idx);
    const po_member* m = &b->layout->m[idx];
    uint32_t d = m->dtype;
    if (!is_scalar(d)) { uint32_t cd = vec3_component(d); uint64_t cs = po_dtype_size(cd, 0); scalar_val s = load_scalar(cd, view_at(&v, p) + comp * cs); return s.cls == 0 ? (double)s.i : s.cls == 1 ? (double)s.u : s.f; }
    scalar_val s = load_scalar(d, view_at(&v, p));
    return s.cls == 0 ? (double)s.i : s.cls == 1 ? (double)s.u : s.f;
}

/* voxel_grid.rs:168-215: max-pool starting from 0.0 */
static double centroid_max_pool(const voxel_t* v, const po_buffer* b, int idx) {
    double cur = 0.0;
    for (uint64_t k = 0; k < v->n; ++k) { double x = attr_as_f64(b, idx, v->pts[k], 0); if (x > cur) cur = x; }
    return cur;
}

/* voxel_grid.rs:218-329: most common value; ties are HashMap-order (nondeterministic) in the reference,
 * the oracle picks the smallest value among the most frequent ones. Integer attributes only. */
static int64_t centroid_most_common(const voxel_t* v, const po_buffer* b, int idx) {
    int64_t best = 0; uint64_t best_count = 0;
    for (uint64_t k = 0; k < v->n; ++k) {
        int64_t val = (int64_t)attr_as_f64(b, idx, v->pts[k], 0);
        uint64_t cnt = 0;
        for (uint64_t j = 0; j < v->n; ++j) if ((int64_t)attr_as_f64(b, idx, v->pts[j], 0) == val) cnt++;
        if (cnt > best_count || (cnt == best_count && val < best)) { best = val; best_count = cnt; }
    }
    return best;
}

/* voxel_grid.rs:333-387 */
static void centroid_average_vec(const voxel_t* v, const po_buffer* b, int idx, double out[3]) {
    double s[3] = {0.0, 0.0, 0.0};
    for (uint64_t k = 0; k < v->n; ++k)
        for (int c = 0; c < 3; ++c) s[c] += attr_as_f64(b, idx, v->pts[k], c);
    double n = (double)v->n;
    for (int c = 0; c < 3; ++c) out[c] = s[c] / n;
}

/* voxel_grid.rs:391-439 */
static double centroid_average_num(const voxel_t* v, const po_buffer* b, int idx) {
    double s = 0.0;
    for (uint64_t k = 0; k < v->n; ++k) s += attr_as_f64(b, idx, v->pts[k], 0);
    return s / (double)v->n;
}

typedef enum { R_POS, R_MEAN_U16, R_MEAN_VEC_U16, R_MEAN_VEC_F32, R_MODE, R_MODE_BOOL, R_MAX_U8, R_MAX_F64, R_MAX_U64 } reduce_kind;
typedef struct { const char* name; uint32_t dtype; reduce_kind kind; } builtin_rule;
/* voxel_grid.rs:461-679, in source order */
static const builtin_rule RULES[] = {
    {"Position3D", PO_VEC3F64, R_POS},
    {"Intensity", PO_U16, R_MEAN_U16},
    {"ReturnNumber", PO_U8, R_MODE},
    {"NumberOfReturns", PO_U8, R_MODE},
    {"ClassificationFlags", PO_U8, R_MAX_U8},
    {"ScannerChannel", PO_U8, R_MODE},
    {"ScanDirectionFlag", PO_U8, R_MODE_BOOL},
    {"EdgeOfFlightLine", PO_U8, R_MODE_BOOL},
    {"Classification", PO_U8, R_MODE},
    {"ScanAngleRank", PO_I8, R_MODE},
    {"ScanAngle", PO_I16, R_MODE},
    {"UserData", PO_U8, R_MODE},
    {"PointSourceID", PO_U16, R_MODE},
    {"ColorRGB", PO_VEC3U16, R_MEAN_VEC_U16},
    {"GpsTime", PO_F64, R_MAX_F64},
    {"NIR", PO_U16, R_MEAN_U16},
    {"PointID", PO_U64, R_MAX_U64},
    {"Normal", PO_VEC3F32, R_MEAN_VEC_F32},
};

static const builtin_rule* find_rule(const po_member* m) {
    for (size_t i = 0; i < sizeof(RULES) / sizeof(RULES[0]); ++i)
        if (strcmp(RULES[i].name, m->name) == 0 && RULES[i].dtype == m->dtype) return &RULES[i];
    return NULL;
}

static void write_target(po_buffer* dst, int tidx, uint64_t row, const void* bytes) {
    raw_view tv = view_attr(dst, tidx);
    memcpy(view_at(&tv, row), bytes, tv.size);
}

typedef struct { uint64_t pos[3]; uint64_t idx; } keyed_pt;
static int cmp_keyed(const void* a, const void* b) {
    const keyed_pt *x = (const keyed_pt*)a, *y = (const keyed_pt*)b;
    int c = cmp_pos(x->pos, y->pos);
    if (c) return c;
    return x->idx < y->idx ? -1 : x->idx > y->idx ? 1 : 0;
}

/* voxel_grid.rs:109-165 */
int po_voxelgrid_filter(const po_buffer* src, double lx, double ly, double lz, po_buffer* dst,
                        uint64_t* voxel_keys, int use_sort) {
    int pi = po_layout_index_of(src->layout, "Position3D", PO_VEC3F64);
    if (pi < 0) return PO_ERR_ATTR_NOT_FOUND; /* :116-121 panic */
    double bmin[3], bmax[3];
    if (po_calculate_bounds(src, bmin, bmax) != 1) return PO_ERR_INVALID; /* :124 unwrap on None */
    /* waveform attrs / non-builtin attrs in the target layout panic (:452-459, :682-687) */
    static const char* WAVE[] = {"WaveformDataOffset", "WaveformPacketSize", "WaveformParameters", "WavePacketDescriptorIndex", "ReturnPointWaveformLocation"};
    static const uint32_t WAVE_T[] = {PO_U64, PO_U32, PO_VEC3F32, PO_U8, PO_F32};
    for (int w = 0; w < 5; ++w) if (po_layout_index_of(dst->layout, WAVE[w], WAVE_T[w]) >= 0) return PO_ERR_UNSUPPORTED;
    int src_idx[PO_MAX_ATTRS];
    const builtin_rule* rules[PO_MAX_ATTRS];
    for (uint32_t a = 0; a < dst->layout->n; ++a) {
        rules[a] = find_rule(&dst->layout->m[a]);
        if (!rules[a]) return PO_ERR_UNSUPPORTED;
        src_idx[a] = po_layout_index_of(src->layout, rules[a]->name, rules[a]->dtype);
        if (src_idx[a] < 0) return PO_ERR_ATTR_NOT_FOUND; /* view_attribute::<T> panics */
    }
    uint64_t nm[3];
    double* mk[3];
    double leaf[3] = {lx, ly, lz};
    for (int c = 0; c < 3; ++c) {
        nm[c] = po_create_markers(bmin[c], bmax[c], leaf[c], NULL, 0);
        mk[c] = (double*)malloc((nm[c] ? nm[c] : 1) * sizeof(double));
        po_create_markers(bmin[c], bmax[c], leaf[c], mk[c], nm[c]);
    }
    raw_view pv = view_attr(src, pi);
    voxel_t* vox = NULL;
    uint64_t nv = 0, capv = 0;
    if (!use_sort) {
        /* faithful: binary_search_by_key + Vec::insert (:141-152) */
        for (uint64_t i = 0; i < src->len; ++i) {
            double p[3];
            memcpy(p, view_at(&pv, i), 24);
            uint64_t pos[3];
            po_find_leaf(p, mk[0], nm[0], mk[1], nm[1], mk[2], nm[2], pos);
            uint64_t lo = 0, hi = nv;
            int found = 0;
            while (lo < hi) {
                uint64_t mid = lo + (hi - lo) / 2;
                int c = cmp_pos(vox[mid].pos, pos);
                if (c == 0) { lo = mid; found = 1; break; }
                if (c < 0) lo = mid + 1; else hi = mid;
            }
            if (found) voxel_push(&vox[lo], i);
            else {
                if (nv == capv) { capv = capv ? capv * 2 : 1024; vox = (voxel_t*)realloc(vox, capv * sizeof(voxel_t)); }
                memmove(&vox[lo + 1], &vox[lo], (nv - lo) * sizeof(voxel_t));
                memset(&vox[lo], 0, sizeof(voxel_t));
                memcpy(vox[lo].pos, pos, sizeof pos);
                voxel_push(&vox[lo], i);
                nv++;
            }
        }
    } else {
        /* equivalent restatement for large inputs: stable sort of (pos, index); validated against the
         * faithful branch in tests/test_oracle_voxel.py */
        keyed_pt* kp = (keyed_pt*)malloc((src->len ? src->len : 1) * sizeof(keyed_pt));
        for (uint64_t i = 0; i < src->len; ++i) {
            double p[3];
            memcpy(p, view_at(&pv, i), 24);
            po_find_leaf_bsearch(p, mk[0], nm[0], mk[1], nm[1], mk[2], nm[2], kp[i].pos);
            kp[i].idx = i;
        }
        qsort(kp, src->len, sizeof(keyed_pt), cmp_keyed);
        for (uint64_t i = 0; i < src->len; ++i) {
            if (i == 0 || cmp_pos(kp[i].pos, kp[i - 1].pos) != 0) {
                if (nv == capv) { capv = capv ? capv * 2 : 1024; vox = (voxel_t*)realloc(vox, capv * sizeof(voxel_t)); }
                memset(&vox[nv], 0, sizeof(voxel_t));
                memcpy(vox[nv].pos, kp[i].pos, sizeof kp[i].pos);
                nv++;
            }
            voxel_push(&vox[nv - 1], kp[i].idx);
        }
        free(kp);
    }
    /* :155-164 */
    for (uint64_t vi = 0; vi < nv; ++vi) {
        const voxel_t* v = &vox[vi];
        if (voxel_keys) memcpy(voxel_keys + 3 * vi, v->pos, 24);
        for (uint32_t a = 0; a < dst->layout->n; ++a) {
            int si = src_idx[a];
            switch (rules[a]->kind) {
                case R_POS: { double c[3]; centroid_average_vec(v, src, si, c); write_target(dst, (int)a, vi, c); break; }
                case R_MEAN_U16: { double m = centroid_average_num(v, src, si); uint16_t r = (uint16_t)f64_as_unsigned(m, 16); write_target(dst, (int)a, vi, &r); break; }
                case R_MEAN_VEC_U16: { double c[3]; centroid_average_vec(v, src, si, c); uint16_t r[3]; for (int k = 0; k < 3; ++k) r[k] = (uint16_t)f64_as_unsigned(c[k], 16); write_target(dst, (int)a, vi, r); break; }
                case R_MEAN_VEC_F32: { double c[3]; centroid_average_vec(v, src, si, c); float r[3]; for (int k = 0; k < 3; ++k) r[k] = (float)c[k]; write_target(dst, (int)a, vi, r); break; }
                case R_MODE: { int64_t m = centroid_most_common(v, src, si); write_target(dst, (int)a, vi, &m); /* `as u8/i8/i16/u16` wraps: low bytes (little endian) */ break; }
                case R_MODE_BOOL: { uint8_t r = centroid_most_common(v, src, si) != 0; write_target(dst, (int)a, vi, &r); break; }
                case R_MAX_U8: { uint8_t r = (uint8_t)f64_as_unsigned(centroid_max_pool(v, src, si), 8); write_target(dst, (int)a, vi, &r); break; }
                case R_MAX_F64: { double r = centroid_max_pool(v, src, si); write_target(dst, (int)a, vi, &r); break; }
                case R_MAX_U64: { uint64_t r = f64_as_unsigned(centroid_max_pool(v, src, si), 64); write_target(dst, (int)a, vi, &r); break; }
            }
        }
    }
    dst->len = nv;
    for (uint64_t vi = 0; vi < nv; ++vi) free(vox[vi].pts);
    free(vox);
    for (int c = 0; c < 3; ++c) free(mk[c]);
    return PO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * K: kNN (brute force stand-in for kd-tree 0.3.0 `nearests`)
 * ---------------------------------------------------------------------------------------------- */

void po_knn_bruteforce(const double* pts, uint64_t n, const double* queries, uint64_t nq, uint32_t k,
                       uint32_t* idx_out, double* d2_out) {
    uint64_t kk = k < n ? k : n;
    for (uint64_t q = 0; q < nq; ++q) {
        uint32_t* bi = idx_out + q * k;
        double* bd = d2_out + q * k;
        uint64_t cnt = 0;
        for (uint64_t i = 0; i < n; ++i) {
            volatile double dx = pts[3 * i] - queries[3 * q], dy = pts[3 * i + 1] - queries[3 * q + 1], dz = pts[3 * i + 2] - queries[3 * q + 2];
            volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
            volatile double s = xx + yy;
            double d2 = s + zz;
            if (cnt == kk && !(d2 < bd[cnt - 1])) continue; /* ties keep the lower index */
            uint64_t pos = cnt < kk ? cnt : kk - 1;
            while (pos > 0 && d2 < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; pos--; }
            bd[pos] = d2;
            bi[pos] = (uint32_t)i;
            if (cnt < kk) cnt++;
        }
        for (uint64_t j = kk; j < k; ++j) { bi[j] = 0xFFFFFFFFu; bd[j] = INFINITY; }
    }
}

/* ------------------------------------------------------------------------------------------------
 * K (second implementation): exact kNN over a kd-tree, the CPU baseline SURVEY 8d asks for ("nanoflann-style,
 * hand-written") and the checker for clouds the O(N^2) brute force cannot reach.  Same contract as
 * po_knn_bruteforce: the k nearest points of the cloud to each query point (the query itself included),
 * ordered by (d2, index); d2 = (dx*dx + dy*dy) + dz*dz in exactly that association, no FMA.  The kd-tree 0.3.0
 * crate the reference calls (normal_estimation.rs:103,108) is not in /root/reference: tie ORDER at exactly equal
 * distances stays unpinned, the neighbour SET and the distances are exact.
 * ---------------------------------------------------------------------------------------------- */

typedef struct { uint32_t lo, hi; int32_t dim; double split; uint32_t left, right; } kd_node; /* dim < 0: leaf [lo,hi) */
typedef struct { const double* pts; uint32_t* perm; kd_node* nodes; uint32_t n_nodes, cap; } kd_tree;
#define KD_LEAF 12

/* quickselect on the unique composite key (coordinate, index): afterwards a[mid] holds the element of rank mid - lo,
 * everything left of it is smaller, everything right of it larger */
static void kd_select(const double* pts, uint32_t* a, uint32_t lo, uint32_t hi, uint32_t mid, int dim) {
#define KD_LESS(x, y) (pts[3 * (size_t)(x) + dim] < pts[3 * (size_t)(y) + dim] || (pts[3 * (size_t)(x) + dim] == pts[3 * (size_t)(y) + dim] && (x) < (y)))
    while (hi - lo > 1) {
        const uint32_t p = a[lo + (hi - lo) / 2];
        uint32_t i = lo, j = hi - 1;
        for (;;) {
            while (KD_LESS(a[i], p)) ++i;
            while (KD_LESS(p, a[j])) --j;
            if (i >= j) break;
            uint32_t t = a[i]; a[i] = a[j]; a[j] = t;
            ++i; --j;
        }
        if (i == j) {          /* both stopped on the pivot: it is in its final place */
            if (mid == i) return;
            if (mid < i) hi = i; else lo = i + 1;
        } else {               /* crossed: [lo, j] < pivot-side, [i, hi) > pivot-side, j + 1 == i */
            if (mid <= j) hi = j + 1; else lo = i;
        }
    }
#undef KD_LESS
}

static uint32_t kd_build(kd_tree* t, uint32_t lo, uint32_t hi) {
    uint32_t id = t->n_nodes++;
    kd_node* nd = &t->nodes[id];
    nd->lo = lo; nd->hi = hi; nd->dim = -1; nd->split = 0.0; nd->left = nd->right = 0;
    if (hi - lo <= KD_LEAF) return id;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = lo; i < hi; ++i)
        for (int c = 0; c < 3; ++c) { double v = t->pts[3 * (size_t)t->perm[i] + c]; if (v < mn[c]) mn[c] = v; if (v > mx[c]) mx[c] = v; }
    int dim = 0;
    for (int c = 1; c < 3; ++c) if (mx[c] - mn[c] > mx[dim] - mn[dim]) dim = c;
    if (!(mx[dim] > mn[dim])) return id; /* all points coincide: stays a (large) leaf */
    uint32_t mid = lo + (hi - lo) / 2;
    kd_select(t->pts, t->perm, lo, hi, mid, dim);
    double split = t->pts[3 * (size_t)t->perm[mid] + dim]; /* left: coordinate <= split, right: >= split */
    uint32_t l = kd_build(t, lo, mid), r = kd_build(t, mid, hi);
    nd = &t->nodes[id]; /* (nodes are preallocated: the pointer is stable, re-read for clarity) */
    nd->dim = dim; nd->split = split; nd->left = l; nd->right = r;
    return id;
}

typedef struct { uint32_t k, cnt; uint32_t* idx; double* d2; } kd_list;

static inline void kd_offer(kd_list* L, double d2, uint32_t i) {
    if (L->cnt == L->k) {
        double wd = L->d2[L->k - 1];
        if (d2 > wd || (d2 == wd && i > L->idx[L->k - 1])) return;
    }
    uint32_t pos = L->cnt < L->k ? L->cnt : L->k - 1;
    while (pos > 0 && (d2 < L->d2[pos - 1] || (d2 == L->d2[pos - 1] && i < L->idx[pos - 1]))) {
        L->d2[pos] = L->d2[pos - 1]; L->idx[pos] = L->idx[pos - 1]; pos--;
    }
    L->d2[pos] = d2; L->idx[pos] = i;
    if (L->cnt < L->k) L->cnt++;
}

static void kd_search(const kd_tree* t, uint32_t id, const double* q, kd_list* L) {
    const kd_node* nd = &t->nodes[id];
    if (nd->dim < 0) {
        for (uint32_t i = nd->lo; i < nd->hi; ++i) {
            uint32_t p = t->perm[i];
            volatile double dx = t->pts[3 * (size_t)p] - q[0], dy = t->pts[3 * (size_t)p + 1] - q[1], dz = t->pts[3 * (size_t)p + 2] - q[2];
            volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
            volatile double s = xx + yy;
            kd_offer(L, s + zz, p);
        }
        return;
    }
    double diff = q[nd->dim] - nd->split;
    uint32_t near = diff <= 0.0 ? nd->left : nd->right, far = diff <= 0.0 ? nd->right : nd->left;
    kd_search(t, near, q, L);
    volatile double dd = diff * diff; /* a lower bound (in the same rounding) of every d2 on the far side */
    if (L->cnt < L->k || !(dd > L->d2[L->k - 1])) kd_search(t, far, q, L); /* equal bound: a tie with a lower index may hide there */
}

typedef struct { const kd_tree* t; uint64_t q0, q1; uint32_t k, kk; uint32_t* idx_out; double* d2_out; uint64_t out_base;
                 double* normals; double* curv; int rc; } kd_job;

static void* kd_worker(void* arg) {
    kd_job* j = (kd_job*)arg;
    uint32_t* li = (uint32_t*)malloc(j->k * sizeof(uint32_t));
    double* ld = (double*)malloc(j->k * sizeof(double));
    double* nb = (double*)malloc(3 * (size_t)j->k * sizeof(double));
    for (uint64_t q = j->q0; q < j->q1; ++q) {
        kd_list L = {j->kk, 0, li, ld};
        kd_search(j->t, 0, j->t->pts + 3 * q, &L);
        if (j->idx_out) {
            uint32_t* oi = j->idx_out + (q - j->out_base) * j->k;
            for (uint32_t s = 0; s < j->k; ++s) oi[s] = s < L.cnt ? li[s] : 0xFFFFFFFFu;
        }
        if (j->d2_out) {
            double* od = j->d2_out + (q - j->out_base) * j->k;
            for (uint32_t s = 0; s < j->k; ++s) od[s] = s < L.cnt ? ld[s] : INFINITY;
        }
        if (j->normals) {
            for (uint32_t s = 0; s < L.cnt; ++s) memcpy(nb + 3 * (size_t)s, j->t->pts + 3 * (size_t)li[s], 24);
            int rc = po_normal_estimation(nb, L.cnt, j->normals + 3 * (q - j->out_base), j->curv + (q - j->out_base));
            if (rc != PO_OK && j->rc == PO_OK) j->rc = rc;
        }
    }
    free(li); free(ld); free(nb);
    return NULL;
}

/* kNN (and optionally normals, normal_estimation.rs:103-127) of the points [q_begin, q_end) of the cloud against the whole
 * cloud.  Outputs are indexed from q_begin.  threads <= 0: one. */
int po_knn_kdtree(const double* pts, uint64_t n, uint64_t q_begin, uint64_t q_end, uint32_t k, uint32_t* idx_out, double* d2_out,
                  double* normals_out, double* curv_out, int threads) {
    if (n == 0 || q_end <= q_begin) return PO_OK;
    if (n > 0xFFFFFFF0ull || q_end > n || k == 0) return PO_ERR_INVALID;
    kd_tree t;
    t.pts = pts;
    t.perm = (uint32_t*)malloc(n * sizeof(uint32_t));
    t.cap = (uint32_t)(2 * (n / (KD_LEAF / 2 + 1) + 2) + 16);
    if (t.cap < 2 * n + 16 && n < 64) t.cap = (uint32_t)(2 * n + 16);
    t.nodes = (kd_node*)malloc((size_t)t.cap * sizeof(kd_node));
    t.n_nodes = 0;
    for (uint64_t i = 0; i < n; ++i) t.perm[i] = (uint32_t)i;
    kd_build(&t, 0, (uint32_t)n);
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > q_end - q_begin) threads = (int)(q_end - q_begin);
    kd_job* jobs = (kd_job*)malloc((size_t)threads * sizeof(kd_job));
    pthread_t* th = (pthread_t*)malloc((size_t)threads * sizeof(pthread_t));
    uint64_t per = (q_end - q_begin + threads - 1) / threads;
    for (int w = 0; w < threads; ++w) {
        kd_job* j = &jobs[w];
        j->t = &t; j->k = k; j->kk = k < n ? k : (uint32_t)n;
        j->q0 = q_begin + per * w; j->q1 = j->q0 + per < q_end ? j->q0 + per : q_end;
        if (j->q0 > q_end) j->q0 = q_end;
        j->idx_out = idx_out; j->d2_out = d2_out; j->out_base = q_begin; j->normals = normals_out; j->curv = curv_out; j->rc = PO_OK;
        pthread_create(&th[w], NULL, kd_worker, j);
    }
    int rc = PO_OK;
    for (int w = 0; w < threads; ++w) { pthread_join(th[w], NULL); if (jobs[w].rc != PO_OK) rc = jobs[w].rc; }
    free(jobs); free(th); free(t.nodes); free(t.perm);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * N: normal estimation (pasture-algorithms/src/normal_estimation.rs)
 * ---------------------------------------------------------------------------------------------- */

static int nb_is_dense(const double* p, uint64_t k) { /* :133-140 */
    for (uint64_t i = 0; i < 3 * k; ++i) if (p[i] != p[i]) return 0;
    return 1;
}
static int pt_is_finite(const double* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); } /* :143-148 */

/* normal_estimation.rs:198-237 */
void po_compute_centroid(const double* pts, uint64_t k, double out[3]) {
    double t[3] = {0.0, 0.0, 0.0};
    if (nb_is_dense(pts, k)) {
        for (uint64_t i = 0; i < k; ++i) { t[0] += pts[3 * i]; t[1] += pts[3 * i + 1]; t[2] += pts[3 * i + 2]; }
        for (int c = 0; c < 3; ++c) out[c] = t[c] / (double)k;
    } else {
        uint64_t cnt = 0;
        for (uint64_t i = 0; i < k; ++i) if (pt_is_finite(pts + 3 * i)) { t[0] += pts[3 * i]; t[1] += pts[3 * i + 1]; t[2] += pts[3 * i + 2]; cnt++; }
        for (int c = 0; c < 3; ++c) out[c] = t[c] / (double)cnt;
    }
}

/* normal_estimation.rs:240-305 ; out9 row-major */
int po_compute_covariance(const double* pts, uint64_t k, double C[9]) {
    for (int i = 0; i < 9; ++i) C[i] = 0.0;
    double cen[3];
    po_compute_centroid(pts, k, cen);
    int dense = nb_is_dense(pts, k);
    uint64_t count = 0;
    for (uint64_t i = 0; i < k; ++i) {
        const double* p = pts + 3 * i;
        if (!dense && !pt_is_finite(p)) continue;
        volatile double d0 = p[0] - cen[0], d1 = p[1] - cen[1], d2 = p[2] - cen[2];
        volatile double m;
        m = d1 * d1; C[4] += m;  /* (1,1) :258 */
        m = d1 * d2; C[5] += m;  /* (1,2) :259 */
        m = d2 * d2; C[8] += m;  /* (2,2) :260 */
        double dx = d0;          /* :262-263 diff_mean *= diff_x */
        volatile double e0 = d0 * dx, e1 = d1 * dx, e2 = d2 * dx;
        C[0] += e0; C[1] += e1; C[2] += e2; /* :265-267 */
        count++;
    }
    if (dense) count = k;
    if (count < 3) return PO_ERR_TOO_FEW_POINTS; /* :296-298 */
    C[3] = C[1]; C[6] = C[2]; C[7] = C[5];       /* :300-302 */
    return PO_OK;
}

/* normal_estimation.rs:308-325 */
static void solve_quadratic(double c2, double c1, double ev[3]) {
    ev[0] = 0.0;
    volatile double a = c2 * c2, b = 4.0 * c1;
    double delta = a - b;
    if (delta < 0.0) delta = 0.0;
    double sd = sqrt(delta);
    ev[2] = 0.5 * (c2 + sd);
    ev[1] = 0.5 * (c2 - sd);
}

/* normal_estimation.rs:328-392. Every product/sum is its own rounding, left-to-right as Rust parses it. */
static void solve_polynomial(const double C[9], double ev[3]) {
#define M(r, c) C[(r) * 3 + (c)]
    volatile double t1 = M(0, 0) * M(1, 1); t1 = t1 * M(2, 2);
    volatile double t2 = 2.0 * M(0, 1); t2 = t2 * M(0, 2); t2 = t2 * M(1, 2);
    volatile double t3 = M(0, 0) * M(1, 2); t3 = t3 * M(1, 2);
    volatile double t4 = M(1, 1) * M(0, 2); t4 = t4 * M(0, 2);
    volatile double t5 = M(2, 2) * M(0, 1); t5 = t5 * M(0, 1);
    volatile double c0 = t1 + t2; c0 = c0 - t3; c0 = c0 - t4; c0 = c0 - t5;
    volatile double u1 = M(0, 0) * M(1, 1), u2 = M(0, 1) * M(0, 1), u3 = M(0, 0) * M(2, 2), u4 = M(0, 2) * M(0, 2), u5 = M(1, 1) * M(2, 2), u6 = M(1, 2) * M(1, 2);
    volatile double c1 = u1 - u2; c1 = c1 + u3; c1 = c1 - u4; c1 = c1 + u5; c1 = c1 - u6;
    volatile double c2 = M(0, 0) + M(1, 1); c2 = c2 + M(2, 2);
#undef M
    if (fabs(c0) < 2.220446049250313e-16) { solve_quadratic(c2, c1, ev); return; } /* :346-347 */
    double one_third = 1.0 / 3.0;
    double sqrt_3 = sqrt(3.0);
    volatile double c2_third = c2 * one_third;
    volatile double a1 = c2 * c2_third;
    volatile double a2 = c1 - a1;
    double alpha_third = a2 * one_third;
    if (alpha_third > 0.0) alpha_third = 0.0;
    volatile double h1 = 2.0 * c2_third; h1 = h1 * c2_third; /* 2.0 * c2_third * c2_third */
    volatile double h2 = h1 - c1;
    volatile double h3 = c2_third * h2;
    volatile double h4 = c0 + h3;
    double half_beta = 0.5 * h4;
    volatile double q1 = half_beta * half_beta;
    volatile double q2 = alpha_third * alpha_third; q2 = q2 * alpha_third;
    double q = q1 + q2;
    if (q > 0.0) q = 0.0;
    double rho = sqrt(-alpha_third);
    volatile double th = atan2(sqrt(-q), half_beta);
    double theta = th * one_third;
    double ct = cos(theta), st = sin(theta);
    volatile double r1 = 2.0 * rho; r1 = r1 * ct;
    ev[0] = c2_third + r1;
    volatile double s3 = sqrt_3 * st;
    volatile double p1 = ct + s3; p1 = rho * p1;
    ev[1] = c2_third - p1;
    volatile double p2 = ct - s3; p2 = rho * p2;
    ev[2] = c2_third - p2;
    /* ascending sort (:381-383) */
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j) if (ev[j + 1] < ev[j]) { double t = ev[j]; ev[j] = ev[j + 1]; ev[j + 1] = t; }
    if (ev[0] <= 0.0) solve_quadratic(c2, c1, ev); /* :386-387 */
}

static void cross3(const double a[3], const double b[3], double r[3]) {
    volatile double m1, m2;
    m1 = a[1] * b[2]; m2 = a[2] * b[1]; r[0] = m1 - m2;
    m1 = a[2] * b[0]; m2 = a[0] * b[2]; r[1] = m1 - m2;
    m1 = a[0] * b[1]; m2 = a[1] * b[0]; r[2] = m1 - m2;
}
static double norm3(const double a[3]) { volatile double x = a[0] * a[0], y = a[1] * a[1], z = a[2] * a[2]; volatile double s = x + y; s = s + z; return sqrt(s); }

/* normal_estimation.rs:395-467 */
void po_solve_plane_parameter(const double C[9], double normal[3], double* curvature) {
    /* eigen_3x3 :429-453 */
    double scale = 0.0; /* covariance_matrix.abs().max() */
    for (int i = 0; i < 9; ++i) { double a = fabs(C[i]); if (a > scale) scale = a; }
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = C[i] / scale;
    double ev[3];
    solve_polynomial(C, ev);            /* eigenvalues of the UNSCALED matrix (:441) */
    double eigen_value = ev[0] * scale; /* :443 */
    /* :446-449 `.diagonal()` returns a copy: the subtraction is a no-op */
    double rows[3][3];
    cross3(&S[0], &S[3], rows[0]); /* :396-400 */
    cross3(&S[0], &S[6], rows[1]);
    cross3(&S[3], &S[6], rows[2]);
    int best = 0;
    for (int r = 0; r < 3; ++r) if (norm3(rows[r]) > norm3(rows[best])) best = r; /* :412-417 strict > */
    for (int c = 0; c < 3; ++c) normal[c] = rows[best][c];
    volatile double es = C[0] + C[4]; es = es + C[8]; /* :459 */
    double eigen_sum = es;
    *curvature = eigen_sum != 0.0 ? fabs(eigen_value / eigen_sum) : 0.0; /* :460-464 */
}

int po_normal_estimation(const double* pts, uint64_t k, double normal[3], double* curvature) {
    double C[9];
    int rc = po_compute_covariance(pts, k, C);
    if (rc) return rc; /* :471 unwrap -> panic */
    po_solve_plane_parameter(C, normal, curvature);
    return PO_OK;
}

/* normal_estimation.rs:79-130 */
int po_compute_normals(const double* pts, uint64_t n, uint32_t k, double* normals_out, double* curv_out) {
    if (n < 3) return PO_ERR_TOO_FEW_POINTS; /* :86-88 */
    if (k < 3) return PO_ERR_INVALID;        /* :89-91 */
    uint32_t* idx = (uint32_t*)malloc(k * sizeof(uint32_t));
    double* d2 = (double*)malloc(k * sizeof(double));
    double* nb = (double*)malloc(3 * (size_t)k * sizeof(double));
    int rc = PO_OK;
    for (uint64_t i = 0; i < n && rc == PO_OK; ++i) {
        po_knn_bruteforce(pts, n, pts + 3 * i, 1, k, idx, d2);
        uint64_t kk = k < n ? k : n;
        for (uint64_t j = 0; j < kk; ++j) memcpy(nb + 3 * j, pts + 3 * (size_t)idx[j], 24);
        rc = po_normal_estimation(nb, kk, normals_out + 3 * i, curv_out + i);
    }
    free(idx); free(d2); free(nb);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * R: reprojection -- closed-form operation pipeline (reprojection.rs:38-45 calls PROJ; the only
 * pinned CRS pair is EPSG:4326 -> EPSG:3309, reprojection.rs:275-289)
 * ---------------------------------------------------------------------------------------------- */

#define PO_PI 3.14159265358979323846

static double albers_q(double e, double sinphi) { /* Snyder 3-12 */
    double es = e * sinphi;
    return (1.0 - e * e) * (sinphi / (1.0 - es * es) - (1.0 / (2.0 * e)) * log((1.0 - es) / (1.0 + es)));
}

void po_reproject(const po_proj_op* ops, uint32_t n_ops, const double* in, double* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        double v[3] = {in[3 * i], in[3 * i + 1], in[3 * i + 2]};
        double z_in = v[2];
        for (uint32_t k = 0; k < n_ops; ++k) {
            const double* p = ops[k].p;
            switch (ops[k].kind) {
                case PO_PROJ_AFFINE: {
                    double r[3];
                    for (int c = 0; c < 3; ++c) r[c] = p[3 * c] * v[0] + p[3 * c + 1] * v[1] + p[3 * c + 2] * v[2] + p[9 + c];
                    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
                    break;
                }
                case PO_PROJ_GEODETIC_TO_ECEF: {
                    double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f);
                    double lat = v[0] * (PO_PI / 180.0), lon = v[1] * (PO_PI / 180.0), h = v[2];
                    double sl = sin(lat), cl = cos(lat);
                    double N = a / sqrt(1.0 - e2 * sl * sl);
                    v[0] = (N + h) * cl * cos(lon);
                    v[1] = (N + h) * cl * sin(lon);
                    v[2] = (N * (1.0 - e2) + h) * sl;
                    break;
                }
                case PO_PROJ_ECEF_TO_GEODETIC: {
                    double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f), b = a * (1.0 - f);
                    double X = v[0], Y = v[1], Z = v[2];
                    double lon = atan2(Y, X);
                    double pr = sqrt(X * X + Y * Y);
                    double lat = atan2(Z, pr * (1.0 - e2));
                    double h = 0.0;
                    for (int it = 0; it < 8; ++it) {
                        double sl = sin(lat);
                        double N = a / sqrt(1.0 - e2 * sl * sl);
                        h = pr / cos(lat) - N;
                        lat = atan2(Z, pr * (1.0 - e2 * N / (N + h)));
                    }
                    (void)b;
                    v[0] = lat; v[1] = lon; v[2] = h;
                    break;
                }
                case PO_PROJ_ALBERS_FWD: {
                    double a = p[0], f = 1.0 / p[1], e2 = f * (2.0 - f), e = sqrt(e2);
                    double phi1 = p[2], phi2 = p[3], phi0 = p[4], lam0 = p[5], x0 = p[6], y0 = p[7];
                    double m1 = cos(phi1) / sqrt(1.0 - e2 * sin(phi1) * sin(phi1));
                    double m2 = cos(phi2) / sqrt(1.0 - e2 * sin(phi2) * sin(phi2));
                    double q0 = albers_q(e, sin(phi0)), q1 = albers_q(e, sin(phi1)), q2 = albers_q(e, sin(phi2));
                    double nn = (m1 * m1 - m2 * m2) / (q2 - q1);
                    double Cc = m1 * m1 + nn * q1;
                    double rho0 = a * sqrt(Cc - nn * q0) / nn;
                    double q = albers_q(e, sin(v[0]));
                    double rho = a * sqrt(Cc - nn * q) / nn;
                    double theta = nn * (v[1] - lam0);
                    v[0] = x0 + rho * sin(theta);
                    v[1] = y0 + rho0 - rho * cos(theta);
                    break;
                }
                case PO_PROJ_SET_Z: v[2] = z_in; break;
                case PO_PROJ_WEBMERC_FWD: {
                    double R = 6378137.0;
                    double lat = v[0] * (PO_PI / 180.0), lon = v[1] * (PO_PI / 180.0);
                    v[0] = R * lon;
                    v[1] = R * log(tan(PO_PI / 4.0 + lat / 2.0));
                    break;
                }
                case PO_PROJ_TMERC_FWD: case PO_PROJ_TMERC_INV: {
                    /* EPSG method 9807 as written in IOGP Guidance Note 7-2 (3.5.3.1, "JHS formulas"); every constant is
                     * recomputed per point from the public parameters a, 1/f, lat0, lon0, k0, FE, FN (radians) */
                    double a = p[0], f = 1.0 / p[1], lat0 = p[2], lon0 = p[3], k0 = p[4], fe = p[5], fn = p[6];
                    double n = f / (2.0 - f), n2 = n * n, n3 = n2 * n, n4 = n2 * n2, e = sqrt(f * (2.0 - f));
                    double B = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0);
                    double h[4] = {n / 2.0 - 2.0 / 3.0 * n2 + 5.0 / 16.0 * n3 + 41.0 / 180.0 * n4,
                                   13.0 / 48.0 * n2 - 3.0 / 5.0 * n3 + 557.0 / 1440.0 * n4,
                                   61.0 / 240.0 * n3 - 103.0 / 140.0 * n4, 49561.0 / 161280.0 * n4};
                    double M0 = 0.0;
                    if (lat0 != 0.0) {
                        double Q0 = asinh(tan(lat0)) - e * atanh(e * sin(lat0));
                        double x0 = atan(sinh(Q0)), x = x0;
                        for (int j = 1; j <= 4; ++j) x += h[j - 1] * sin(2.0 * j * x0);
                        M0 = B * x;
                    }
                    if (ops[k].kind == PO_PROJ_TMERC_FWD) {
                        double Q = asinh(tan(v[0])) - e * atanh(e * sin(v[0]));
                        double beta = atan(sinh(Q));
                        double eta0 = atanh(cos(beta) * sin(v[1] - lon0));
                        double xi0 = asin(sin(beta) * cosh(eta0));
                        double xi = xi0, eta = eta0;
                        for (int j = 1; j <= 4; ++j) {
                            xi += h[j - 1] * sin(2.0 * j * xi0) * cosh(2.0 * j * eta0);
                            eta += h[j - 1] * cos(2.0 * j * xi0) * sinh(2.0 * j * eta0);
                        }
                        v[0] = fe + k0 * B * eta;
                        v[1] = fn + k0 * (B * xi - M0);
                    } else {
                        double hi[4] = {n / 2.0 - 2.0 / 3.0 * n2 + 37.0 / 96.0 * n3 - 1.0 / 360.0 * n4,
                                        1.0 / 48.0 * n2 + 1.0 / 15.0 * n3 - 437.0 / 1440.0 * n4,
                                        17.0 / 480.0 * n3 - 37.0 / 840.0 * n4, 4397.0 / 161280.0 * n4};
                        double eta = (v[0] - fe) / (B * k0), xi = ((v[1] - fn) + k0 * M0) / (B * k0);
                        double xi0 = xi, eta0 = eta;
                        for (int j = 1; j <= 4; ++j) {
                            xi0 -= hi[j - 1] * sin(2.0 * j * xi) * cosh(2.0 * j * eta);
                            eta0 -= hi[j - 1] * cos(2.0 * j * xi) * sinh(2.0 * j * eta);
                        }
                        double beta = asin(sin(xi0) / cosh(eta0));
                        double Qp = asinh(tan(beta)), Q = Qp;
                        for (int it = 0; it < 12; ++it) Q = Qp + e * atanh(e * tanh(Q));
                        v[0] = atan(sinh(Q));
                        v[1] = lon0 + asin(tanh(eta0) / cos(beta));
                    }
                    break;
                }
                case PO_PROJ_DEG2RAD_LATLON: v[0] *= PO_PI / 180.0; v[1] *= PO_PI / 180.0; break;
                case PO_PROJ_RAD2DEG_LATLON: v[0] *= 180.0 / PO_PI; v[1] *= 180.0 / PO_PI; break;
                case PO_PROJ_WEBMERC_INV: { /* GN7-2 3.5.1.2 reverse: D = -N/a, lat = pi/2 - 2 atan(e^D) */
                    double R = 6378137.0;
                    double lon = v[0] / R, lat = PO_PI / 2.0 - 2.0 * atan(exp(-v[1] / R));
                    v[0] = lat * (180.0 / PO_PI); v[1] = lon * (180.0 / PO_PI);
                    break;
                }
                default: break;
            }
        }
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}

uint32_t po_pipeline_epsg4326_to_3309(po_proj_op* ops) {
    memset(ops, 0, 5 * sizeof(po_proj_op));
    ops[0].kind = PO_PROJ_GEODETIC_TO_ECEF; ops[0].p[0] = 6378137.0; ops[0].p[1] = 298.257223563; /* WGS84 */
    ops[1].kind = PO_PROJ_AFFINE; /* inverse of NAD27->WGS84 towgs84=-8,160(159),176(175): +8, -159, -175 */
    ops[1].p[0] = 1.0; ops[1].p[4] = 1.0; ops[1].p[8] = 1.0; ops[1].p[9] = 8.0; ops[1].p[10] = -159.0; ops[1].p[11] = -175.0;
    ops[2].kind = PO_PROJ_ECEF_TO_GEODETIC; ops[2].p[0] = 6378206.4; ops[2].p[1] = 294.978698213898; /* Clarke 1866 */
    ops[3].kind = PO_PROJ_ALBERS_FWD; ops[3].p[0] = 6378206.4; ops[3].p[1] = 294.978698213898;
    ops[3].p[2] = 34.0 * PO_PI / 180.0; ops[3].p[3] = 40.5 * PO_PI / 180.0; ops[3].p[4] = 0.0;
    ops[3].p[5] = -120.0 * PO_PI / 180.0; ops[3].p[6] = 0.0; ops[3].p[7] = -4000000.0;
    ops[4].kind = PO_PROJ_SET_Z;
    return 5;
}

/* ------------------------------------------------------------------------------------------------
 * RANSAC segmentation, pasture-algorithms/src/segmentation.rs
 * ---------------------------------------------------------------------------------------------- */

void po_ransac_model_from_samples(int kind, const double* pts, const uint64_t* s, double* m) {
    if (kind == 0) { /* generate_rng_plane :60-76 */
        const double* pa = pts + 3 * s[0]; const double* pb = pts + 3 * s[1]; const double* pc = pts + 3 * s[2];
        double v1[3], v2[3], nrm[3];
        for (int c = 0; c < 3; ++c) { v1[c] = pb[c] - pa[c]; v2[c] = pc[c] - pa[c]; }
        nrm[0] = v1[1] * v2[2] - v1[2] * v2[1]; /* nalgebra cross */
        nrm[1] = v1[2] * v2[0] - v1[0] * v2[2];
        nrm[2] = v1[0] * v2[1] - v1[1] * v2[0];
        double dot = (nrm[0] * pa[0] + nrm[1] * pa[1]) + nrm[2] * pa[2]; /* nalgebra dot, 3 rows: a + b + c */
        m[0] = nrm[0]; m[1] = nrm[1]; m[2] = nrm[2]; m[3] = -dot;
    } else { /* generate_rng_line :91-95 */
        for (int c = 0; c < 3; ++c) { m[c] = pts[3 * s[0] + c]; m[3 + c] = pts[3 * s[1] + c]; }
    }
}

double po_ransac_distance(int kind, const double* m, const double p[3]) {
    if (kind == 0) { /* distance_point_plane :31-35 */
        double d = fabs(((m[0] * p[0] + m[1] * p[1]) + m[2] * p[2]) + m[3]);
        double e = sqrt((m[0] * m[0] + m[1] * m[1]) + m[2] * m[2]);
        return d / e;
    }
    /* distance_point_line :39-44 */
    double v[3], w[3], cr[3];
    for (int c = 0; c < 3; ++c) { v[c] = m[3 + c] - m[c]; w[c] = m[c] - p[c]; }
    cr[0] = v[1] * w[2] - v[2] * w[1];
    cr[1] = v[2] * w[0] - v[0] * w[2];
    cr[2] = v[0] * w[1] - v[1] * w[0];
    double num = sqrt((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
    double den = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    return num / den;
}

void po_ransac_rank_models(int kind, const double* pts, uint64_t n, const double* models, uint64_t n_models,
                           double threshold, uint64_t* rankings) {
    const int w = kind == 0 ? 4 : 6;
    for (uint64_t h = 0; h < n_models; ++h) { /* generate_plane_model :119-138 / generate_line_model :98-117 */
        uint64_t r = 0;
        for (uint64_t i = 0; i < n; ++i)
            if (po_ransac_distance(kind, models + w * h, pts + 3 * i) < threshold) r++;
        rankings[h] = r;
    }
}

uint64_t po_ransac_inliers(int kind, const double* pts, uint64_t n, const double* model, double threshold, uint64_t* indices) {
    uint64_t r = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (po_ransac_distance(kind, model, pts + 3 * i) < threshold) indices[r++] = i;
    return r;
}

void po_ransac_draw_samples(int kind, uint64_t n, uint64_t n_models, uint64_t seed, uint64_t* samples) {
    uint64_t j = 0;
    for (uint64_t h = 0; h < n_models; ++h) {
        uint64_t r1 = po_splitmix64(seed, j++) % n;
        uint64_t r2 = po_splitmix64(seed, j++) % n;
        while (r1 == r2) r2 = po_splitmix64(seed, j++) % n; /* :52-54, :85-89 */
        if (kind == 0) {
            uint64_t r3 = po_splitmix64(seed, j++) % n;
            while (r2 == r3 || r1 == r3) r3 = po_splitmix64(seed, j++) % n; /* :55-59 */
            samples[3 * h] = r1; samples[3 * h + 1] = r2; samples[3 * h + 2] = r3;
        } else {
            samples[2 * h] = r1; samples[2 * h + 1] = r2;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic inputs: identical streams on CPU and GPU
 * ---------------------------------------------------------------------------------------------- */

uint64_t po_splitmix64(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

/* C2 stream: raw LAS fmt-0 records (20 B) */
void po_gen_las_fmt0_records(uint8_t* out, uint64_t first, uint64_t n, uint64_t seed) {
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        uint8_t* r = out + 20 * k;
        for (int c = 0; c < 3; ++c) {
            int32_t v = (int32_t)(po_splitmix64(seed, 8 * i + (uint64_t)c) % 2000001ull) - 1000000;
            memcpy(r + 4 * c, &v, 4);
        }
        uint64_t h3 = po_splitmix64(seed, 8 * i + 3), h4 = po_splitmix64(seed, 8 * i + 4), h5 = po_splitmix64(seed, 8 * i + 5);
        uint16_t inten = (uint16_t)(h3 >> 48);
        memcpy(r + 12, &inten, 2);
        r[14] = (uint8_t)(h4);
        r[15] = (uint8_t)(h4 >> 8);
        r[16] = (uint8_t)(h4 >> 16);
        r[17] = (uint8_t)(h4 >> 24);
        uint16_t psid = (uint16_t)h5;
        memcpy(r + 18, &psid, 2);
    }
}

/* C1 stream: LasPointFormat0 default layout (35 B packed) */
void po_gen_c1_points(uint8_t* out, uint64_t first, uint64_t n, uint64_t seed, const double offset[3]) {
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        uint8_t* r = out + 35 * k;
        for (int c = 0; c < 3; ++c) {
            volatile double m = u01(po_splitmix64(seed, 8 * i + (uint64_t)c)) * 2000.0;
            volatile double s = m - 1000.0;
            double v = s + offset[c];
            memcpy(r + 8 * c, &v, 8);
        }
        uint64_t h3 = po_splitmix64(seed, 8 * i + 3), h4 = po_splitmix64(seed, 8 * i + 4), h5 = po_splitmix64(seed, 8 * i + 5);
        uint16_t inten = (uint16_t)(h3 >> 48);
        memcpy(r + 24, &inten, 2);
        r[26] = (uint8_t)(h4 & 7);
        r[27] = (uint8_t)((h4 >> 3) & 7);
        r[28] = (uint8_t)((h4 >> 6) & 1);
        r[29] = (uint8_t)((h4 >> 7) & 1);
        r[30] = (uint8_t)(h4 >> 8);
        r[31] = (uint8_t)(h4 >> 16);
        r[32] = (uint8_t)(h4 >> 24);
        uint16_t psid = (uint16_t)h5;
        memcpy(r + 33, &psid, 2);
    }
}

/* C3/C4 stream: terrain-like positions built from + - * only */
void po_gen_terrain_positions(double* out, uint64_t first, uint64_t n, uint64_t seed) {
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        volatile double x = u01(po_splitmix64(seed, 8 * i + 0)) * 500.0;
        volatile double y = u01(po_splitmix64(seed, 8 * i + 1)) * 500.0;
        volatile double a = x * 0.002, b = y * 0.002;
        volatile double aa = a * a, bb = b * b, ab = a * b;
        volatile double t1 = aa - bb; t1 = 10.0 * t1;
        volatile double t2 = 5.0 * ab;
        volatile double t3 = 0.1 * u01(po_splitmix64(seed, 8 * i + 2));
        volatile double z = t1 + t2; z = z + t3;
        out[3 * k] = x; out[3 * k + 1] = y; out[3 * k + 2] = z;
    }
}
