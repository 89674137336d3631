"""GPU parity of the RANSAC segmentation (pasture-algorithms/src/segmentation.rs) against the oracle: same draws ->
same models bit for bit, same rankings, same inlier index lists; plus the reference's own test assertions."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import HashMapBuffer, VectorBuffer
from pasture_b200 import algorithms as A
from tests.test_oracle_segmentation import setup_point_cloud

pytestmark = pytest.mark.gpu


def cloud(pts, columnar=True, device="cuda", extra=False):
    attrs = [("Position3D", pb.PointAttributeDataType.Vec3f64)]
    if extra:
        attrs = [("Intensity", pb.PointAttributeDataType.U16), ("GpsTime", pb.PointAttributeDataType.F64)] + attrs
    layout = pb.PointLayout.from_attributes([pb.PointAttributeDefinition.custom(n, t) for n, t in attrs])
    buf = (HashMapBuffer if columnar else VectorBuffer)(layout, len(pts), device)
    buf.set_attribute("Position3D", np.ascontiguousarray(pts, dtype=np.float64))
    return buf


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("columnar,device,extra", [(True, "cuda", False), (False, "cuda", True), (True, "cpu", False), (False, "cpu", True)])
def test_reference_cloud_same_draws(kind, columnar, device, extra):
    pts = setup_point_cloud()
    buf = cloud(pts, columnar, device, extra)
    samples = O.ransac_draw_samples(kind, len(pts), 300, 11)
    models, ranks = A.ransac_rank_samples(buf, kind, samples, 0.1)
    want_models = O.ransac_models(kind, pts, samples)
    assert np.array_equal(models.view(np.uint64), want_models.view(np.uint64))
    assert np.array_equal(ranks, O.ransac_rank_models(kind, pts, want_models, 0.1))
    best = int(np.argmax(ranks))
    got = A.ransac_inliers(buf, kind, models[best], 0.1).cpu().numpy()
    assert np.array_equal(got.astype(np.uint64), O.ransac_inliers(kind, pts, want_models[best], 0.1))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_ransac_plane_like_the_reference_tests(seed):  # segmentation.rs:417-441
    buf = cloud(setup_point_cloud())
    plane, idx = A.ransac_plane_par(buf, 0.1, 300, seed=seed)
    assert plane.ranking == idx.numel() == 1600
    got = set(idx.cpu().tolist())
    assert all(i in got for i in range(2000) if i % 5 != 3)
    m, r, widx = O.ransac(0, setup_point_cloud(), 0.1, 300, seed)
    assert [plane.a, plane.b, plane.c, plane.d] == m.tolist() and r == plane.ranking
    assert np.array_equal(idx.cpu().numpy().astype(np.uint64), widx)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_ransac_line_like_the_reference_tests(seed):  # :443-457
    buf = cloud(setup_point_cloud())
    line, idx = A.ransac_line_serial(buf, 0.1, 300, seed=seed)
    assert line.ranking == idx.numel() == 400
    got = set(idx.cpu().tolist())
    assert all(i in got for i in range(2000) if i % 5 == 3)
    m, r, widx = O.ransac(1, setup_point_cloud(), 0.1, 300, seed)
    assert np.array_equal(np.concatenate([line.first, line.second]), m) and r == line.ranking


@pytest.mark.parametrize("kind", [0, 1])
def test_threshold_edge_cases_are_decided_like_the_reference_expression(kind):
    """points placed exactly on, one ulp inside and one ulp outside of the threshold distance, awkward thresholds,
    degenerate models (zero normal / zero-length line), NaN and inf coordinates"""
    rng = np.random.default_rng(5)
    n = 20000
    pts = rng.uniform(-50, 50, (n, 3))
    pts[:2000, 2] = 0.0  # in the plane z = 0 / spread around
    models = []
    if kind == 0:
        models += [[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 3.0, -1.5], [1.0, 2.0, 2.0, 0.25], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                   [1e-200, 0.0, 0.0, 0.0], [1e200, 1e200, 0.0, 0.0]]
    else:
        models += [[0.0, 0.0, 0.0, 1.0, 0.0, 0.0], [1.0, 1.0, 1.0, 4.0, 5.0, 13.0], [2.0, 2.0, 2.0, 2.0, 2.0, 2.0],
                   [0.0, 0.0, 0.0, 1e-180, 0.0, 0.0], [0.0, 0.0, 0.0, 1e160, 0.0, 0.0]]
    models = np.array(models)
    # distances of a few points to model 0/1 are used as thresholds: forces `distance == threshold` ties and near-ties
    thr_list = [0.5, 1e-3, 0.0, -1.0, float("nan"), float("inf"), 1e-300, 5e-324]
    for j in (0, 7, 2500, 9000):
        d = O.lib().po_ransac_distance(kind, O._ptr(models[1].copy()), O._ptr(pts[j].copy()))
        thr_list += [d, np.nextafter(d, np.inf), np.nextafter(d, -np.inf)]
    pts[100] = [np.nan, 1.0, 2.0]
    pts[101] = [np.inf, 1.0, 2.0]
    pts[102] = [1e300, -1e300, 1e300]
    buf = cloud(pts)
    for thr in thr_list:
        got = A.ransac_rank_models(buf, kind, models, thr)
        want = O.ransac_rank_models(kind, pts, models, thr)
        assert np.array_equal(got, want), thr
        for m in models[:3]:
            gi = A.ransac_inliers(buf, kind, m, thr).cpu().numpy().astype(np.uint64)
            assert np.array_equal(gi, O.ransac_inliers(kind, pts, m, thr)), thr


def test_many_models_and_ragged_sizes():
    for n in (3, 5, 1023, 1025, 4099):
        pts = O.gen_terrain_positions(0, n)
        buf = cloud(pts, columnar=False, extra=True)
        for kind in (0, 1):
            samples = O.ransac_draw_samples(kind, n, 700, 3)  # 3 launches of 256 models
            models, ranks = A.ransac_rank_samples(buf, kind, samples, 0.75)
            assert np.array_equal(ranks, O.ransac_rank_models(kind, pts, O.ransac_models(kind, pts, samples), 0.75))


def test_contract():
    buf2 = cloud(np.zeros((2, 3)))
    with pytest.raises(pb.PastureB200Error) as e:  # "buffer needs to include at least 3 points to generate a plane."
        A.ransac_plane_par(buf2, 0.1, 10, seed=1)
    assert e.value.code == -9
    with pytest.raises(pb.PastureB200Error) as e:  # "... at least 2 points to generate a line."
        A.ransac_line_par(cloud(np.zeros((1, 3))), 0.1, 10, seed=1)
    assert e.value.code == -9
    layout = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D.with_custom_datatype(pb.PointAttributeDataType.Vec3f32)])
    with pytest.raises(pb.PastureB200Error) as e:  # view_attribute::<Vector3<f64>> panics on a Vec3f32 position
        A.ransac_plane_par(HashMapBuffer(layout, 10, "cuda"), 0.1, 10, seed=1)
    assert e.value.code == -1
    with pytest.raises(pb.PastureB200Error):
        A.ransac_rank_samples(cloud(np.zeros((5, 3))), 0, np.array([[0, 1, 9]], dtype=np.uint64), 0.1)


def test_full_size_terrain_plane():
    """20 M terrain points, 300 models: the rankings of a sample of models agree with the oracle on a subsample scaled
    check (exact on the first 200k points), and the winner's index list is consistent with its ranking"""
    n = 20_000_000
    buf = A.synth_terrain_positions(n)
    samples = O.ransac_draw_samples(0, n, 300, 21)
    models, ranks = A.ransac_rank_samples(buf, 0, samples, 0.5)
    best = len(ranks) - 1 - int(np.argmax(ranks[::-1]))
    idx = A.ransac_inliers(buf, 0, models[best], 0.5)
    assert idx.numel() == int(ranks[best])
    assert bool((idx[1:] > idx[:-1]).all())
    # exact check of all 300 models on a prefix, through a prefix view of the same buffer
    m = 200_000
    pts = O.gen_terrain_positions(0, m)
    pre = cloud(pts)
    assert np.array_equal(A.ransac_rank_models(pre, 0, models, 0.5), O.ransac_rank_models(0, pts, models, 0.5))
    plane, idx2 = A.ransac_plane_par(buf, 0.5, 300, seed=21)
    assert plane.ranking == int(ranks[best]) and torch.equal(idx2, idx)
