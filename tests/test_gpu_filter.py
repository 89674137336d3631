"""GPU parity of HashMapBuffer::filter / filter_into (pasture-core/src/containers/point_buffer.rs:1064-1136,
test :2296-2329) against the same selection done with numpy boolean indexing."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import HashMapBuffer, VectorBuffer
from tests import util

pytestmark = pytest.mark.gpu
BIG = [("GpsTime", O.F64), ("ColorRGB", O.VEC3U16), ("Position3D", O.VEC3F64), ("Classification", O.U8), ("Intensity", O.I16)]


@pytest.mark.parametrize("src_col", [True, False])
@pytest.mark.parametrize("dst_type", [HashMapBuffer, VectorBuffer])
@pytest.mark.parametrize("device", ["cuda", "cpu"])
@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("n,keep", [(1, 1.0), (17, 0.5), (5000, 0.3), (100003, 0.9), (4096, 0.0), (8192, 1.0)])
def test_filter_matches_boolean_indexing(src_col, dst_type, device, packed, n, keep):
    ol, pl = util.layouts(BIG, packed=packed)
    ob, pbuf = util.random_bytes_buffers(ol, pl, n, src_col, seed=n, device=device)
    rng = np.random.default_rng(n + 1)
    mask = rng.random(n) < keep
    out = pb.filter(pbuf, mask, dst_type)
    assert out.len() == int(mask.sum())
    want = O.OBuffer(ol, int(mask.sum()), dst_type is HashMapBuffer)
    assert O.filter_into(ob, lambda i: bool(mask[i]), want) == out.len()
    for i in range(len(pl)):
        assert np.array_equal(out._attribute_bytes(i), want.attribute_bytes(i)), pl.at(i)


def test_filter_every_second_point_like_the_reference_test():  # point_buffer.rs:2296-2329
    ol, pl = util.layouts(BIG, packed=1)
    ob, pbuf = util.random_bytes_buffers(ol, pl, 16, True, seed=3)
    mask = np.arange(16) % 2 == 0
    even = pb.filter(pbuf, mask, VectorBuffer)
    assert even.len() == 8
    assert np.array_equal(even.view_attribute("Position3D"), ob.attribute("Position3D")[::2])


def test_filter_into_contract():
    ol, pl = util.layouts(BIG, packed=1)
    _, pbuf = util.random_bytes_buffers(ol, pl, 100, True, seed=4)
    mask = np.ones(100, bool)
    with pytest.raises(pb.PastureB200Error) as e:  # "buffer.len() must be at least as large as the number of predicate matches"
        pb.filter_into(pbuf, HashMapBuffer(pl, 50, "cuda"), mask)
    assert e.value.code == -5
    _, other = util.layouts(BIG)
    with pytest.raises(pb.PastureB200Error) as e:  # "PointLayouts must match"
        pb.filter_into(pbuf, HashMapBuffer(other, 100, "cuda"), mask)
    assert e.value.code == -4
    # larger target: only the first num_matches points are written, padding of interleaved targets survives
    dst = VectorBuffer(pl, 120, "cuda")
    dst.data.fill_(0xEE)
    mask[::3] = False
    k = pb.filter_into(pbuf, dst, mask)
    assert k == int(mask.sum())
    raw = dst.raw_bytes().reshape(120, pl.size_of_point_entry())
    assert np.all(raw[k:] == 0xEE)
    # a layout with padding: the padding bytes of the written records are not touched either (the reference copies
    # attribute by attribute)
    ol2, pl2 = util.layouts(BIG)
    assert pl2.size_of_point_entry() > sum(m.size() for m in pl2.attributes())
    _, src2 = util.random_bytes_buffers(ol2, pl2, 100, False, seed=5)
    dst2 = VectorBuffer(pl2, 100, "cuda")
    dst2.data.fill_(0xEE)
    k2 = pb.filter_into(src2, dst2, mask)
    raw2 = dst2.raw_bytes().reshape(100, pl2.size_of_point_entry())
    covered = np.zeros(pl2.size_of_point_entry(), bool)
    for m in pl2.attributes():
        covered[m.offset():m.offset() + m.size()] = True
    assert np.all(raw2[:k2][:, ~covered] == 0xEE) and not np.all(raw2[:k2][:, covered] == 0xEE)


def test_filter_full_size():
    """100 M points of the C2 stream: keep the points of one class; count and a checksum of a column agree with torch"""
    n = 100_000_000
    src = pb.algorithms.synth_las_fmt0_records(n)
    raw = pb.PointLayout.las_raw(0)
    rec = src.data[: 20 * n].view(n, 20)
    mask = (rec[:, 15] & 3) == 1  # "classification % 4 == 1"
    out = pb.filter(src, mask, VectorBuffer)
    torch.cuda.synchronize()
    k = int(mask.sum().item())
    assert out.len() == k
    got = out.data[: 20 * k].view(k, 20)
    assert bool(((got[:, 15] & 3) == 1).all())
    assert int(got[:, 12].to(torch.int64).sum().item()) == int(rec[:, 12][mask].to(torch.int64).sum().item())
    sel = torch.nonzero(mask)[:1000, 0]
    assert torch.equal(got[:1000], rec[sel])
