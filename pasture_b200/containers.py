"""Point buffers -- host-side mirror of pasture-core/src/containers/point_buffer.rs.

VectorBuffer (:659, interleaved / AoS) and HashMapBuffer (:1031, columnar / SoA, one contiguous byte
array per attribute, Vec3 values stay packed xyz).  Storage is a torch uint8 tensor per array, in host
(optionally pinned) or device memory: PyTorch is only the allocator here, the library sees raw pointers
through pb200_buffer_desc, exactly like ExternalMemoryBuffer (:1479) wraps foreign memory.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import BufferDesc
from .layout import DT, PointAttributeDefinition, PointLayout

_NP = {DT.U8: np.uint8, DT.I8: np.int8, DT.U16: np.uint16, DT.I16: np.int16, DT.U32: np.uint32, DT.I32: np.int32,
       DT.U64: np.uint64, DT.I64: np.int64, DT.F32: np.float32, DT.F64: np.float64}
_VEC3 = {DT.Vec3u8: DT.U8, DT.Vec3u16: DT.U16, DT.Vec3f32: DT.F32, DT.Vec3i32: DT.I32, DT.Vec3f64: DT.F64}

INTERLEAVED, COLUMNAR = 0, 1
HOST, DEVICE = 0, 1


def _alloc(nbytes, device, pinned, uninitialized=False):
    dev = torch.device(device)
    make = torch.empty if uninitialized else torch.zeros
    if dev.type == "cpu":
        return make(max(1, nbytes), dtype=torch.uint8, pin_memory=bool(pinned) and torch.cuda.is_available())
    return make(max(1, nbytes), dtype=torch.uint8, device=dev)


def _typed(raw_bytes, dtype, n):
    """(n, size) uint8 numpy -> typed numpy array"""
    raw = np.ascontiguousarray(raw_bytes)
    if dtype in _NP:
        return raw.view(_NP[dtype]).reshape(n)
    if dtype in _VEC3:
        return raw.view(_NP[_VEC3[dtype]]).reshape(n, 3)
    return raw


def _untyped(values, dtype, n, size):
    comp = _NP.get(dtype) or _NP.get(_VEC3.get(dtype))
    if comp is None:
        return np.ascontiguousarray(values, dtype=np.uint8).reshape(n, size)
    return np.ascontiguousarray(np.asarray(values, dtype=comp)).view(np.uint8).reshape(n, size)


class _BufferBase:
    def len(self):
        return self._len

    def __len__(self):
        return self._len

    def is_empty(self):
        return self._len == 0

    def point_layout(self):
        return self._layout

    @property
    def device(self):
        return self._device

    def _memspace(self):
        return DEVICE if self._device.type == "cuda" else HOST

    def slice_first(self, attribute):
        """typed value of the attribute in point 0 (one small copy; the seed of minmax.rs' fold)"""
        idx = self._layout.index_of(attribute)
        m = self._layout.at(idx)
        if hasattr(self, "columns"):
            raw = self.columns[idx][: m.size()]
        else:
            raw = self.data[m.offset(): m.offset() + m.size()]
        return _typed(raw.cpu().numpy().reshape(1, m.size()), m.datatype(), 1)[0]

    def view_attribute(self, attribute):
        """typed host copy of one attribute (AttributeView, buffer_views.rs:291). name + datatype must match."""
        if isinstance(attribute, PointAttributeDefinition):
            idx = self._layout.index_of(attribute)
            if idx is None:
                raise KeyError(f"attribute {attribute} not in PointLayout")  # view_attribute panics
        else:
            idx = self._layout.index_by_name(attribute)
            if idx is None:
                raise KeyError(attribute)
        m = self._layout.at(idx)
        return _typed(self._attribute_bytes(idx), m.datatype(), self._len)


class VectorBuffer(_BufferBase):
    """Interleaved buffer: one byte array of len * size_of_point_entry (point_buffer.rs:659-945)."""

    def __init__(self, layout, length=0, device="cpu", pinned=False, data=None, uninitialized=False):
        self._layout = layout
        self._len = int(length)
        self._device = torch.device(device)
        self._pinned = pinned
        if data is not None:
            self.data = data
            self._device = data.device
        else:  # resize(): zero fill :833-837 (uninitialized: the caller promises to write every byte, e.g. convert())
            self.data = _alloc(self._len * layout.size_of_point_entry(), device, pinned, uninitialized)

    @classmethod
    def new_from_layout(cls, layout, device="cpu"):
        return cls(layout, 0, device)

    @classmethod
    def with_capacity(cls, capacity, layout, device="cpu"):
        return cls(layout, 0, device)

    @classmethod
    def from_bytes(cls, layout, raw, device="cpu", pinned=False):
        raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
        size = layout.size_of_point_entry()
        assert size > 0 and raw.size % size == 0
        b = cls(layout, raw.size // size, "cpu", pinned)
        b.data[: raw.size] = torch.from_numpy(raw.copy())
        return b.to(device) if torch.device(device).type != "cpu" else b

    def resize(self, n):
        old = self.data
        size = self._layout.size_of_point_entry()
        self.data = _alloc(n * size, self._device, self._pinned)
        keep = min(n, self._len) * size
        self.data[:keep] = old[:keep]
        self._len = int(n)

    def as_interleaved(self):
        return self

    def as_columnar(self):
        return None

    def to(self, device, pinned=False):
        dev = torch.device(device)
        if dev.type == "cpu" and pinned:
            data = _alloc(self.data.numel(), "cpu", True)
            data.copy_(self.data)
        else:
            data = self.data.to(dev)
        return VectorBuffer(self._layout, self._len, dev, pinned, data=data)

    def get_point_range_ref(self, start, end):
        size = self._layout.size_of_point_entry()
        return self.data[start * size: end * size]

    def _attribute_bytes(self, idx):
        m = self._layout.at(idx)
        size = self._layout.size_of_point_entry()
        rec = self.data[: self._len * size].cpu().numpy().reshape(self._len, size)
        return rec[:, m.offset(): m.offset() + m.size()]

    def set_attribute(self, attribute, values):
        idx = self._layout.index_of(attribute) if isinstance(attribute, PointAttributeDefinition) \
            else self._layout.index_by_name(attribute)
        m = self._layout.at(idx)
        size = self._layout.size_of_point_entry()
        host = self.data[: self._len * size].cpu().numpy().reshape(self._len, size).copy()
        host[:, m.offset(): m.offset() + m.size()] = _untyped(values, m.datatype(), self._len, m.size())
        self.data[: self._len * size] = torch.from_numpy(host.reshape(-1)).to(self.data.device)

    def raw_bytes(self):
        return self.data[: self._len * self._layout.size_of_point_entry()].cpu().numpy()

    def desc(self):
        d = BufferDesc()
        d.layout = self._layout._h
        d.kind = INTERLEAVED
        d.memspace = self._memspace()
        d.len = self._len
        d.aos = self.data.data_ptr()
        d.columns = None
        self._keep = (d,)
        return d


class HashMapBuffer(_BufferBase):
    """Columnar buffer: one byte array per attribute in layout order (point_buffer.rs:1031-1474)."""

    def __init__(self, layout, length=0, device="cpu", pinned=False, columns=None, uninitialized=False):
        self._layout = layout
        self._len = int(length)
        self._device = torch.device(device)
        self._pinned = pinned
        if columns is not None:
            self.columns = columns
            if columns:
                self._device = columns[0].device
        else:
            self.columns = [_alloc(self._len * a.size(), device, pinned, uninitialized) for a in layout.attributes()]

    @classmethod
    def new_from_layout(cls, layout, device="cpu"):
        return cls(layout, 0, device)

    @classmethod
    def with_capacity(cls, capacity, layout, device="cpu"):
        return cls(layout, 0, device)

    def resize(self, n):
        new_cols = []
        for col, a in zip(self.columns, self._layout.attributes()):
            c = _alloc(n * a.size(), self._device, self._pinned)
            keep = min(n, self._len) * a.size()
            c[:keep] = col[:keep]
            new_cols.append(c)
        self.columns = new_cols
        self._len = int(n)

    def as_interleaved(self):
        return None

    def as_columnar(self):
        return self

    def to(self, device, pinned=False):
        dev = torch.device(device)
        cols = []
        for c in self.columns:
            if dev.type == "cpu" and pinned:
                t = _alloc(c.numel(), "cpu", True)
                t.copy_(c)
            else:
                t = c.to(dev)
            cols.append(t)
        return HashMapBuffer(self._layout, self._len, dev, pinned, columns=cols)

    def get_attribute_range_ref(self, attribute, start, end):
        idx = self._layout.index_of(attribute)
        size = self._layout.at(idx).size()
        return self.columns[idx][start * size: end * size]

    def _attribute_bytes(self, idx):
        size = self._layout.at(idx).size()
        return self.columns[idx][: self._len * size].cpu().numpy().reshape(self._len, size)

    def set_attribute(self, attribute, values):
        idx = self._layout.index_of(attribute) if isinstance(attribute, PointAttributeDefinition) \
            else self._layout.index_by_name(attribute)
        m = self._layout.at(idx)
        raw = _untyped(values, m.datatype(), self._len, m.size()).reshape(-1)
        self.columns[idx][: raw.size] = torch.from_numpy(raw.copy()).to(self.columns[idx].device)

    def desc(self):
        d = BufferDesc()
        d.layout = self._layout._h
        d.kind = COLUMNAR
        d.memspace = self._memspace()
        d.len = self._len
        d.aos = None
        ptrs = (C.c_void_p * max(1, len(self.columns)))(*[c.data_ptr() for c in self.columns])
        d.columns = C.cast(ptrs, C.POINTER(C.c_void_p))
        self._keep = (d, ptrs)
        return d


def _mask_tensor(mask, n, device):
    m = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(mask)))
    m = (m != 0).to(torch.uint8).reshape(-1)
    assert m.numel() == n, "one predicate value per point"
    return m.to(device).contiguous()


def filter_into(source, target, mask, ctx=None):
    """HashMapBuffer::filter_into (point_buffer.rs:1086): the predicate `Fn(usize) -> bool` is given as a mask over
    the point indices. Returns the number of matches."""
    from ._lib import check, lib
    from .context import context_for, get_context
    ctx = context_for(ctx, source)
    m = _mask_tensor(mask, source.len(), source.device)
    sd, dd = source.desc(), target.desc()
    count = C.c_uint64(0)
    check(lib().pb200_filter_into(ctx._h, C.byref(sd), C.c_void_p(m.data_ptr()), C.byref(dd), C.byref(count)))
    return int(count.value)


def filter(source, mask, out_buffer_type=None, device=None, ctx=None):
    """HashMapBuffer::filter (point_buffer.rs:1064): new buffer holding the points whose mask entry is non-zero"""
    m = _mask_tensor(mask, source.len(), source.device)
    n = int(m.sum().item())
    out_type = out_buffer_type or type(source)
    target = out_type(source.point_layout(), n, device if device is not None else source.device)
    filter_into(source, target, m, ctx)
    return target


def buffers_equal(a, b):
    """attribute-wise byte equality of two buffers with the same attributes (any memory layout)"""
    la, lb = a.point_layout(), b.point_layout()
    if len(la) != len(lb) or a.len() != b.len():
        return False
    for i in range(len(la)):
        if not np.array_equal(a._attribute_bytes(i), b._attribute_bytes(i)):
            return False
    return True
