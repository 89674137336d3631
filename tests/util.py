"""Shared helpers: build the same layouts / buffers for the oracle (oracle/) and the product (pasture_b200)."""
import numpy as np
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import PointAttributeDefinition, PointLayout, FieldAlignment


def layouts(attrs, packed=0):
    """attrs: list of (name, dtype code[, extra_size]) -> (OLayout, PointLayout) built by the same rule"""
    ol = O.OLayout.from_attributes(attrs, packed=packed)
    pl = PointLayout()
    for a in attrs:
        extra = a[2] if len(a) > 2 else 0
        pl.add_attribute(PointAttributeDefinition(a[0], a[1], extra), FieldAlignment(packed))
    assert_same_layout(ol, pl)
    return ol, pl


def assert_same_layout(ol, pl):
    assert ol.size == pl.size_of_point_entry() and ol.align == pl.alignment()
    got = [(m.name(), int(m.datatype()), m.offset(), m.size()) for m in pl.attributes()]
    assert got == ol.members(), (got, ol.members())


def las_layouts(fmt, raw):
    ol = O.OLayout.las_raw(fmt) if raw else O.OLayout.las_default(fmt)
    pl = PointLayout.las_raw(fmt) if raw else PointLayout.las_default(fmt)
    assert_same_layout(ol, pl)
    return ol, pl


def random_bytes_buffers(ol, pl, n, columnar, seed, device="cuda", finite_floats=True):
    """the same pseudo-random content in an oracle buffer and a pasture_b200 buffer"""
    rng = np.random.default_rng(seed)
    ob = O.OBuffer(ol, n, columnar)
    for i, (name, dtype, off, sz) in enumerate(ol.members()):
        if dtype in O.NP_DTYPES or dtype in O.VEC3_COMPONENT:
            comp = O.NP_DTYPES.get(dtype) or O.NP_DTYPES[O.VEC3_COMPONENT[dtype]]
            shape = (n,) if dtype in O.NP_DTYPES else (n, 3)
            if np.issubdtype(comp, np.floating):
                if finite_floats:
                    vals = ((rng.random(shape) - 0.5) * 10.0 ** rng.integers(-3, 12, shape)).astype(comp)
                else:
                    vals = rng.integers(0, 256, shape + (np.dtype(comp).itemsize,), dtype=np.uint8).view(comp).reshape(shape)
            else:
                info = np.iinfo(comp)
                vals = rng.integers(info.min, info.max, shape, dtype=comp, endpoint=True)
            if n:
                ob.set_attribute(name, vals)
        elif n:
            raw = rng.integers(0, 256, (n, sz), dtype=np.uint8)
            if columnar:
                ob.columns[i][: n * sz] = raw.reshape(-1)
            else:
                ob.aos[: n * ol.size].reshape(n, ol.size)[:, off:off + sz] = raw
    return ob, to_pb(ob, pl, device)


def to_pb(ob, pl, device="cuda", pinned=False):
    """copy an oracle buffer into a pasture_b200 buffer of the same memory layout"""
    if ob.columnar:
        b = pb.HashMapBuffer(pl, ob.len, "cpu", pinned)
        for i, a in enumerate(pl.attributes()):
            nb = ob.len * a.size()
            b.columns[i][:nb] = torch.from_numpy(ob.columns[i][:nb].copy())
    else:
        b = pb.VectorBuffer(pl, ob.len, "cpu", pinned)
        nb = ob.len * pl.size_of_point_entry()
        b.data[:nb] = torch.from_numpy(ob.aos[:nb].copy())
    return b.to(device) if torch.device(device).type != "cpu" else b


def assert_buffers_match(ob, pbuf, what=""):
    """every attribute of the product buffer is byte-identical to the oracle buffer"""
    assert ob.len == pbuf.len(), (what, ob.len, pbuf.len())
    for i, (name, dtype, off, sz) in enumerate(ol_members(ob)):
        a = ob.attribute_bytes(i)
        b = pbuf._attribute_bytes(i)
        if not np.array_equal(a, b):
            bad = np.nonzero(np.any(a != b, axis=1))[0]
            raise AssertionError(f"{what}: attribute {name} differs at {len(bad)} points, first {bad[:5]}: "
                                 f"oracle {a[bad[0]].tolist()} gpu {b[bad[0]].tolist()}")


def ol_members(ob):
    return ob.layout.members()


def oracle_transform(t):
    """pasture_b200.Transform -> oracle Transform struct"""
    return O.make_transform(t.kind, s=t.s, o=t.o, shift=t.shift, mask=t.mask)
