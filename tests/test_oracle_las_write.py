"""Pins the oracle's restatement of the default-layout LAS writer (raw_writers.rs:203-362) through the reference's
round-trip property (pasture-io/tests/las_io.rs:245-350): write -> read gives the points back, for every format."""
import numpy as np
import pytest

import oracle as O


def random_default_points(fmt, n, seed):
    """tests/common/mod.rs:56-78 style: integer coordinates in +-1000 so that scale 0.001 round-trips exactly"""
    rng = np.random.default_rng(seed)
    l = O.OLayout.las_default(fmt)
    b = O.OBuffer(l, n, False)
    ext = fmt >= 6
    for name, dtype, off, sz in l.members():
        if name == "Position3D":
            b.set_attribute(name, rng.integers(-1000, 1000, (n, 3)).astype(np.float64))
        elif name in ("ReturnNumber", "NumberOfReturns"):
            b.set_attribute(name, rng.integers(0, 16 if ext else 8, n))
        elif name == "ClassificationFlags":
            b.set_attribute(name, rng.integers(0, 16, n))
        elif name == "ScannerChannel":
            b.set_attribute(name, rng.integers(0, 4, n))
        elif name in ("ScanDirectionFlag", "EdgeOfFlightLine"):
            b.set_attribute(name, rng.integers(0, 2, n))
        elif dtype in O.NP_DTYPES and np.issubdtype(O.NP_DTYPES[dtype], np.integer):
            info = np.iinfo(O.NP_DTYPES[dtype])
            b.set_attribute(name, rng.integers(info.min, info.max, n, dtype=O.NP_DTYPES[dtype], endpoint=True))
        elif dtype == O.VEC3U16:
            b.set_attribute(name, rng.integers(0, 65536, (n, 3)))
        elif dtype == O.VEC3F32:
            b.set_attribute(name, rng.random((n, 3)).astype(np.float32))
        else:
            b.set_attribute(name, rng.random(n) * 1000)
    return l, b


@pytest.mark.parametrize("fmt", range(11))
def test_write_then_read_roundtrip(fmt):
    l, src = random_default_points(fmt, 333, fmt)
    scale, offset = (0.001, 0.001, 0.001), (0.0, 0.0, 0.0)
    rec, counts, mn, mx, panics = O.las_write_points(src, fmt, scale, offset)
    assert panics == 0 and rec.shape == (333, O.OLayout.las_raw(fmt).size)
    raw = O.OLayout.las_raw(fmt)
    rb = O.OBuffer(raw, 333, False)
    rb.aos[:] = rec.reshape(-1)
    back = O.OConverter.las_default(raw, l, scale, offset).convert(rb, False)
    for i in range(l.n):
        assert np.array_equal(back.attribute_bytes(i), src.attribute_bytes(i)), l.members()[i]
    rn = src.attribute("ReturnNumber")
    for r in range(1, 16):
        assert counts[r] == np.count_nonzero(rn == r)
    pos = src.attribute("Position3D")
    assert np.array_equal(mn, pos.min(0)) and np.array_equal(mx, pos.max(0))


def test_write_masks_and_panics():
    l, src = random_default_points(0, 8, 1)
    src.set_attribute("ReturnNumber", [0, 1, 7, 8, 9, 15, 200, 255])  # masked with 0b111 on write (write_helpers.rs:32)
    src.set_attribute("NumberOfReturns", [7] * 8)
    pos = src.attribute("Position3D").copy()
    pos[3] = [3e6, 0, 0]
    src.set_attribute("Position3D", pos)
    rec, counts, mn, mx, panics = O.las_write_points(src, 0, (0.001,) * 3, (0.0,) * 3)
    assert panics == 1
    flags = rec[:, 14]
    assert list(flags & 7) == [0, 1, 7, 0, 1, 7, 0, 7] and np.all(((flags >> 3) & 7) == 7)
    assert counts[1] == 1 and counts[7] == 1 and counts[8] == 1 and counts[15] == 1 and counts[0] == 0
