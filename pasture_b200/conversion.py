"""BufferLayoutConverter -- host-side mirror of pasture-core/src/layout/conversion/buffer_conversion.rs:98-359
and of get_default_las_converter (pasture-io/src/las/raw_readers.rs:31-167).

Arbitrary closures cannot cross the FFI to the GPU, so `set_custom_mapping_with_transformation` takes one of
the enumerated transforms below (the closures that exist in-tree)."""
import ctypes as C

from ._lib import Transform as _CTransform
from ._lib import check, lib
from .containers import HashMapBuffer, VectorBuffer
from .context import context_for, get_context

T_NONE, T_SCALE_OFFSET, T_INV_SCALE_OFFSET, T_ADD, T_SHIFT_MASK = range(5)


class Transform:
    def __init__(self, kind, s=(1.0, 1.0, 1.0), o=(0.0, 0.0, 0.0), shift=0, mask=0, datatype=None):
        self.kind, self.s, self.o, self.shift, self.mask, self.datatype = kind, tuple(s), tuple(o), shift, mask, datatype

    def _c(self):
        t = _CTransform()
        t.kind, t.shift, t.mask = self.kind, self.shift, self.mask
        for i in range(3):
            t.s[i] = float(self.s[i])
            t.o[i] = float(self.o[i])
        return t


def _v3(x):
    return (x, x, x) if isinstance(x, (int, float)) else tuple(x)


def ScaleOffset(scale, offset, datatype=None):
    """|v| v * scale + offset (raw_readers.rs:42-55)"""
    return Transform(T_SCALE_OFFSET, s=_v3(scale), o=_v3(offset), datatype=datatype)


def InvScaleOffset(scale, offset, datatype=None):
    """|v| (v - offset) / scale (write_helpers.rs:15-17)"""
    return Transform(T_INV_SCALE_OFFSET, s=_v3(scale), o=_v3(offset), datatype=datatype)


def Add(offset, datatype=None):
    """|v| v + offset (pnts_reader.rs:265-277, buffer_conversion.rs:780-782)"""
    return Transform(T_ADD, o=_v3(offset), datatype=datatype)


def ShiftMask(shift, mask, datatype=None):
    """|v| (v >> shift) & mask (raw_readers.rs:61-164)"""
    return Transform(T_SHIFT_MASK, shift=shift, mask=mask, datatype=datatype)


class BufferLayoutConverter:
    def __init__(self, handle, from_layout, to_layout, ctx):
        self._h, self._from, self._to, self._ctx = handle, from_layout, to_layout, ctx

    def __del__(self):
        try:
            if self._h:
                lib().pb200_converter_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def _create(cls, from_layout, to_layout, with_default, ctx):
        # a converter belongs to ONE device (its context); pass ctx=get_context(device) for buffers on another GPU
        ctx = ctx or get_context()
        h = C.c_void_p()
        check(lib().pb200_converter_create(ctx._h, from_layout._h, to_layout._h, 1 if with_default else 0, C.byref(h)))
        return cls(h, from_layout, to_layout, ctx)

    @classmethod
    def for_layouts(cls, from_layout, to_layout, ctx=None):  # :112
        return cls._create(from_layout, to_layout, False, ctx)

    @classmethod
    def for_layouts_with_default(cls, from_layout, to_layout, ctx=None):  # :126
        return cls._create(from_layout, to_layout, True, ctx)

    def set_custom_mapping(self, from_attribute, to_attribute):  # :156
        check(lib().pb200_converter_set_custom_mapping(self._h, from_attribute.name().encode(),
                                                       int(from_attribute.datatype()), to_attribute.name().encode(),
                                                       int(to_attribute.datatype())))

    def set_custom_mapping_with_transformation(self, from_attribute, to_attribute, transform, apply_to_source_attribute):  # :194
        tdt = transform.datatype
        if tdt is None:
            tdt = from_attribute.datatype() if apply_to_source_attribute else to_attribute.datatype()
        t = transform._c()
        check(lib().pb200_converter_set_custom_mapping_with_transformation(
            self._h, from_attribute.name().encode(), int(from_attribute.datatype()), to_attribute.name().encode(),
            int(to_attribute.datatype()), int(tdt), C.byref(t), 1 if apply_to_source_attribute else 0))

    def num_mappings(self):
        return lib().pb200_converter_num_mappings(self._h)

    def convert(self, source_buffer, out_buffer_type, device=None):  # :242
        """new_from_layout + resize + convert_into.  The library writes EVERY byte of the new buffer (zeros where no
        mapping applies), so the target is allocated uninitialised and partially mapped interleaved records need no
        read-modify-write"""
        dev = device if device is not None else source_buffer.device
        n = source_buffer.len()
        target = out_buffer_type(self._to, n, dev, uninitialized=True)
        context_for(self._ctx, source_buffer, target)
        sd, dd = source_buffer.desc(), target.desc()
        check(lib().pb200_converter_convert_fresh_range(self._h, C.byref(sd), 0, n, C.byref(dd), 0, n, None))
        return target

    def convert_into_fresh(self, source_buffer, target_buffer, source_range=None, target_range=None):
        """`convert` semantics on an existing buffer (pb200_converter_convert_fresh_range): every byte of the target range is
        written -- zeros where no mapping applies -- and nothing of its previous content is read"""
        context_for(self._ctx, source_buffer, target_buffer)
        sr = source_range if source_range is not None else range(0, source_buffer.len())
        tr = target_range if target_range is not None else range(0, len(sr))
        sd, dd = source_buffer.desc(), target_buffer.desc()
        check(lib().pb200_converter_convert_fresh_range(self._h, C.byref(sd), sr.start, sr.stop, C.byref(dd), tr.start, tr.stop, None))

    def convert_into(self, source_buffer, target_buffer, count_out_of_range=False):  # :268
        return self.convert_into_range(source_buffer, range(0, source_buffer.len()), target_buffer,
                                       range(0, source_buffer.len()), count_out_of_range)

    def convert_into_range(self, source_buffer, source_range, target_buffer, target_range, count_out_of_range=False):  # :292
        context_for(self._ctx, source_buffer, target_buffer)
        sd, dd = source_buffer.desc(), target_buffer.desc()
        oor = C.c_uint64(0)
        check(lib().pb200_converter_convert_into_range(self._h, C.byref(sd), source_range.start, source_range.stop,
                                                       C.byref(dd), target_range.start, target_range.stop,
                                                       C.byref(oor) if count_out_of_range else None))
        return int(oor.value) if count_out_of_range else None

    def convert_into_range_with_bounds(self, source_buffer, source_range, target_buffer, target_range):
        """convert_into_range fused with calculate_bounds over the produced POSITION_3D; returns (min, max) or None"""
        context_for(self._ctx, source_buffer, target_buffer)
        sd, dd = source_buffer.desc(), target_buffer.desc()
        mn, mx, some = (C.c_double * 3)(), (C.c_double * 3)(), C.c_int(0)
        check(lib().pb200_converter_convert_into_range_with_bounds(
            self._h, C.byref(sd), source_range.start, source_range.stop, C.byref(dd), target_range.start,
            target_range.stop, mn, mx, C.byref(some)))
        return (tuple(mn), tuple(mx)) if some.value else None

    def convert_into_range_with_bounds_device(self, source_buffer, source_range, target_buffer, target_range, minmax6):
        """as above, but leaves [min xyz, -max xyz] in the CUDA tensor `minmax6` (6 x f64) without synchronising"""
        context_for(self._ctx, source_buffer, target_buffer)
        sd, dd = source_buffer.desc(), target_buffer.desc()
        return check(lib().pb200_converter_convert_into_range_with_bounds_device(
            self._h, C.byref(sd), source_range.start, source_range.stop, C.byref(dd), target_range.start,
            target_range.stop, C.c_void_p(minmax6.data_ptr())))

    def convert_into_range_with_global_bounds(self, source_buffer, source_range, target_buffer, target_range, comm, minmax6):
        """collective over `comm` (sharding.PeerComm): converts this rank's range and leaves the bounds of ALL ranks'
        produced POSITION_3D as [min xyz, -max xyz] in the CUDA tensor `minmax6` -- one kernel, no NCCL call: the
        exchange happens over peer memory in the kernel's last CTA"""
        context_for(self._ctx, source_buffer, target_buffer)
        sd, dd = source_buffer.desc(), target_buffer.desc()
        return check(lib().pb200_converter_convert_into_range_with_global_bounds(
            self._h, C.byref(sd), source_range.start, source_range.stop, C.byref(dd), target_range.start,
            target_range.stop, comm._h, C.c_void_p(minmax6.data_ptr())))


def get_default_las_converter(raw_las_layout, target_layout, scale, offset, ctx=None):
    """raw_readers.rs:31-167 (las_header.transforms() -> scale/offset per axis)"""
    ctx = ctx or get_context()
    h = C.c_void_p()
    s, o = (C.c_double * 3)(*scale), (C.c_double * 3)(*offset)
    check(lib().pb200_las_default_converter(ctx._h, raw_las_layout._h, target_layout._h, s, o, C.byref(h)))
    return BufferLayoutConverter(h, raw_las_layout, target_layout, ctx)


def transform_attribute(buffer, attribute, transform, ctx=None):
    """BorrowedMutBuffer::transform_attribute (point_buffer.rs:391-404), in place"""
    ctx = context_for(ctx, buffer)
    d = buffer.desc()
    t = transform._c()
    check(lib().pb200_transform_attribute(ctx._h, C.byref(d), attribute.name().encode(), int(attribute.datatype()), C.byref(t)))


def view_attribute_with_conversion(buffer, attribute, ctx=None):
    """AttributeViewConverting (buffer_views.rs:533-650): attribute `attribute.name()` materialised as
    `attribute.datatype()`; returns a typed numpy array"""
    import numpy as np
    import torch

    from .containers import _typed
    ctx = context_for(ctx, buffer)
    n, size = buffer.len(), attribute.size()
    out = torch.zeros(max(1, n * size), dtype=torch.uint8, device=buffer.device)
    d = buffer.desc()
    check(lib().pb200_view_attribute_with_conversion(ctx._h, C.byref(d), attribute.name().encode(),
                                                     int(attribute.datatype()), C.c_void_p(out.data_ptr())))
    if buffer.device.type == "cuda":
        torch.cuda.synchronize(buffer.device)
    return _typed(out[: n * size].cpu().numpy().reshape(n, size), attribute.datatype(), n)
