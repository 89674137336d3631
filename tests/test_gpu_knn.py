"""GPU parity of kNN / radius search over the LBVH (K10-K12), normal estimation (K13) and reprojection (K14).

kNN: neighbour indices AND squared distances bit-exact vs the brute-force oracle, ordered by (d2, index)
(tie order of the reference's kd-tree crate is unpinned).  Normals: bit-exact given identical
neighbour order (no FMA); curvature within 1e-9 relative (device atan2/cos/sin vs glibc).  Reprojection: the 4
reference KATs at the reference's tolerance (1e-4 m), and 1e-6 m vs the oracle elsewhere."""
import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import HashMapBuffer, VectorBuffer, attributes as A
from pasture_b200.algorithms import (compute_normals, knn, radius_search, reproject_point_cloud_between,
                                     reproject_point_cloud_within)
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["register_list", "smem_heap"], autouse=True)
def knn_list_variant(request):
    """every test of this module runs with both k-list implementations (registers / shared-memory heap)"""
    ctx = pb.get_context()
    ctx.set_param("knn.heap", 1 if request.param == "smem_heap" else 0)
    yield request.param
    ctx.set_param("knn.heap", 1)  # the default


def cloud(pts, columnar=True, device="cuda", extra=False):
    attrs = [("Position3D", O.VEC3F64)] + ([("Intensity", O.U16)] if extra else [])
    ol, pl = util.layouts(attrs)
    ob = O.OBuffer(ol, len(pts), columnar)
    if len(pts):
        ob.set_attribute("Position3D", pts)
    return util.to_pb(ob, pl, device)


def clouds():
    rng = np.random.default_rng(7)
    yield "uniform", rng.random((4000, 3)) * [10, 10, 2]
    yield "terrain", O.gen_terrain_positions(0, 6000)
    yield "clustered", np.concatenate([rng.normal(c, 0.05, (700, 3)) for c in ([0, 0, 0], [5, 5, 5], [5, 0, 1], [-3, 2, 0])])
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), np.arange(6.0)), -1).reshape(-1, 3)
    yield "lattice_with_ties", g  # many exactly equal distances
    yield "duplicates", np.repeat(rng.random((300, 3)), 4, axis=0)
    yield "line", np.stack([np.linspace(0, 1, 1000), np.zeros(1000), np.zeros(1000)], 1)


@pytest.mark.parametrize("name,pts", list(clouds()))
@pytest.mark.parametrize("k", [1, 3, 16])
def test_knn_matches_bruteforce(name, pts, k):
    oidx, od2 = O.knn_bruteforce(pts, pts, k)
    idx, d2 = knn(cloud(pts), k)
    torch.cuda.synchronize()
    idx = idx.cpu().numpy().view(np.uint32)
    d2 = d2.cpu().numpy()
    assert np.array_equal(d2, od2), name
    assert np.array_equal(idx, oidx), name


@pytest.mark.parametrize("k", [5, 17, 32, 33, 64])
def test_knn_every_list_width(k):
    """k = 4/16/32 use register-resident lists of that width, k > 32 a local-memory list; bucket boundaries (8 points)
    and partially filled last buckets are hit by the odd sizes"""
    rng = np.random.default_rng(k)
    for n in (k + 1, 8 * k + 3, 3001):
        pts = rng.random((n, 3)) * [4, 4, 1]
        oidx, od2 = O.knn_bruteforce(pts, pts, k)
        idx, d2 = knn(cloud(pts), k)
        assert np.array_equal(d2.cpu().numpy(), od2), (k, n)
        assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx), (k, n)


def test_knn_larger_cloud_deep_tree():
    """30 k terrain points (3750 buckets): the traversal, not the priming range, finds most neighbours"""
    pts = O.gen_terrain_positions(0, 30000)
    oidx, od2 = O.knn_bruteforce(pts, pts, 16)
    idx, d2 = knn(cloud(pts), 16)
    assert np.array_equal(d2.cpu().numpy(), od2)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)


def test_knn_one_million_points_full_parity_tie_order_unpinned():
    """1 M points of the C4 stream, k = 16: EVERY neighbour index and distance against the oracle's exact kd-tree
    (itself checked against the brute force in the CPU suite).  The order inside exact distance ties is (d2, index) on
    both sides; the reference's kd-tree 0.3.0 is not in /root/reference, so that tie order is unpinned."""
    n = 1_000_000
    pts = O.gen_terrain_positions(0, n)
    oidx, od2 = O.knn_kdtree(pts, 16)
    src = pb.algorithms.synth_terrain_positions(n)
    idx, d2 = knn(src, 16)
    assert torch.equal(d2.cpu(), torch.from_numpy(od2))
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)
    # normals of the same cloud: bit-exact normals, curvature within 1e-9 relative (device atan2/cos vs glibc)
    on, oc = O.compute_normals_kdtree(pts, 16)
    nrm, curv = compute_normals(src, 16)
    assert np.array_equal(nrm.cpu().numpy(), on)
    np.testing.assert_allclose(curv.cpu().numpy(), oc, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n,first,count", [(50_000, 0, 50_000), (50_000, 12_345, 6_000), (50_000, 49_990, 10), (50_000, 7, 0),
                                           (400_000, 300_000, 100_000), (1_000_000, 125_000, 125_000)])
def test_query_ranges_equal_the_slices_of_the_full_result(n, first, count):
    """pb200_knn_range / pb200_compute_normals_range (the replicas-only multi-GPU cut: a rank answers the points of its
    own range against the whole cloud): exactly the rows [first, first + count) of the full result, for ranges at the
    start, in the middle, at the very end, empty, and 1/8 of a 1 M-point cloud (queries spread over the Morton order)"""
    pts = O.gen_terrain_positions(0, n)
    oidx, od2 = O.knn_kdtree(pts, 16, first, first + count)
    src = pb.algorithms.synth_terrain_positions(n)
    r = range(first, first + count)
    idx, d2 = knn(src, 16, query_range=r)
    assert idx.shape[0] == count
    assert np.array_equal(d2.cpu().numpy(), od2) and np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)
    nrm, curv = compute_normals(src, 16, query_range=r)
    on, oc = O.compute_normals_kdtree(pts, 16, first, first + count)
    assert np.array_equal(nrm.cpu().numpy(), on)
    np.testing.assert_allclose(curv.cpu().numpy(), oc, rtol=1e-9, atol=1e-12)
    with pytest.raises(pb.PastureB200Error) as e:
        knn(src, 16, query_range=range(n - 5, n + 1))
    assert e.value.code == -5


def test_knn_georeferenced_coordinates():
    """UTM-like magnitudes (5.4e6 m) with decimetre spacing: boxes are f32 intervals relative to the AABB minimum and
    must stay conservative (result bit-exact) -- and a cloud whose points differ only in the last f64 bits"""
    rng = np.random.default_rng(21)
    pts = rng.random((5000, 3)) * [40.0, 40.0, 3.0] + [500000.0, 5400000.0, 100.0]
    oidx, od2 = O.knn_bruteforce(pts, pts, 16)
    idx, d2 = knn(cloud(pts), 16)
    assert np.array_equal(d2.cpu().numpy(), od2) and np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)
    tiny = np.array([5400000.0, 5400000.0, 5400000.0]) + rng.integers(0, 64, (600, 3)) * 2.0 ** -30
    oidx, od2 = O.knn_bruteforce(tiny, tiny, 16)
    idx, d2 = knn(cloud(tiny), 16)
    assert np.array_equal(d2.cpu().numpy(), od2) and np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)


def test_knn_all_points_identical():
    """every code and every distance ties: order must be by original index"""
    pts = np.tile(np.array([[1.5, -2.0, 3.25]]), (100, 1))
    idx, d2 = knn(cloud(pts), 16)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), np.tile(np.arange(16, dtype=np.uint32), (100, 1)))
    assert not d2.cpu().numpy().any()


@pytest.mark.parametrize("columnar,device", [(True, "cuda"), (False, "cuda"), (True, "cpu"), (False, "cpu")])
def test_knn_buffer_kinds_and_small_inputs(columnar, device):
    rng = np.random.default_rng(3)
    pts = rng.random((1500, 3))
    oidx, od2 = O.knn_bruteforce(pts, pts, 8)
    idx, d2 = knn(cloud(pts, columnar, device, extra=True), 8)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx) and np.array_equal(d2.cpu().numpy(), od2)
    for n in (1, 2, 5):  # fewer points than k: the tail is 0xFFFFFFFF / inf
        p = rng.random((n, 3))
        idx, d2 = knn(cloud(p, columnar, device), 8)
        oidx, od2 = O.knn_bruteforce(p, p, 8)
        assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx) and np.array_equal(d2.cpu().numpy(), od2)


def test_radius_search():
    rng = np.random.default_rng(11)
    pts = rng.random((3000, 3))
    r, m = 0.08, 32
    idx, cnt = radius_search(cloud(pts), r, m)
    idx = idx.cpu().numpy().view(np.uint32)
    cnt = cnt.cpu().numpy()
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2)
    d2 = (d2[..., 0] + d2[..., 1]) + d2[..., 2]
    for i in range(0, 3000, 37):
        inside = np.nonzero(d2[i] <= r * r)[0]
        order = inside[np.lexsort((inside, d2[i][inside]))][:m]
        assert cnt[i] == len(order)
        assert np.array_equal(idx[i, : cnt[i]], order)


KAT = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])


def test_compute_normal_kat():  # normal_estimation.rs:580-610
    normals, curv = compute_normals(cloud(KAT, False, extra=True), 3)
    normals, curv = normals.cpu().numpy(), curv.cpu().numpy()
    assert np.all(normals[:, 0] == 0) and np.all(normals[:, 1] == 0) and np.all(normals[:, 2] != 0)
    assert np.all(curv == 0)
    on, oc = O.compute_normals(KAT, 3)
    assert np.array_equal(normals, on) and np.array_equal(curv, oc)


def test_compute_normal_panics():  # :612-698
    with pytest.raises(pb.PastureB200Error) as e:
        compute_normals(cloud(KAT[:1]), 3)
    assert e.value.code == -9
    with pytest.raises(pb.PastureB200Error):
        compute_normals(cloud(KAT[:2]), 3)
    for k in (1, 2):
        with pytest.raises(pb.PastureB200Error) as e:
            compute_normals(cloud(KAT), k)
        assert e.value.code == -8


@pytest.mark.parametrize("name,pts", [c for c in clouds() if c[0] in ("uniform", "terrain", "clustered")])
@pytest.mark.parametrize("k", [3, 16])
def test_normals_match_oracle(name, pts, k):
    pts = pts[:2500]
    on, oc = O.compute_normals(pts, k)
    normals, curv = compute_normals(cloud(pts), k)
    normals, curv = normals.cpu().numpy(), curv.cpu().numpy()
    assert np.array_equal(normals, on), name           # bit-exact: same neighbour order, no FMA
    # curvature = |lambda0 * maxabs(C) / trace(C)| goes through atan2/cos/sin: 1e-9 relative (device libm vs glibc);
    # for rank-deficient neighbourhoods (k = 3, planar patches) lambda0 is pure cancellation noise ~1e-16 * trace,
    # hence the absolute floor
    assert np.allclose(curv, oc, rtol=1e-9, atol=1e-12 * max(1.0, float(np.abs(pts).max()) ** 2)), (name, np.abs(curv - oc).max())


def test_normals_c4_shape_properties():
    """C4 stream at 2 M points: normals of the smooth terrain point mostly upwards, curvature finite and >= 0"""
    n = 2_000_000
    src = pb.algorithms.synth_terrain_positions(n)
    normals, curv = compute_normals(src, 16)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(normals).all()) and bool(torch.isfinite(curv).all()) and bool((curv >= 0).all())
    nz = normals[:, 2].abs() / normals.norm(dim=1).clamp_min(1e-300)
    assert float((nz > 0.5).double().mean()) > 0.9
    # spot-check 64 points against the brute-force oracle on the same cloud
    pts = src.view_attribute(A.POSITION_3D)
    sel = np.linspace(0, n - 1, 64).astype(np.int64)
    oidx, od2 = O.knn_bruteforce(pts, pts[sel], 16)
    for row, i in enumerate(sel):
        nrm = np.zeros(3)
        c = np.zeros(1)
        O.lib().po_normal_estimation(O._ptr(np.ascontiguousarray(pts[oidx[row]])), 16, O._ptr(nrm), O._ptr(c))
        assert np.array_equal(normals[i].cpu().numpy(), nrm)
        assert np.isclose(float(curv[i]), c[0], rtol=1e-9, atol=1e-12 * 500.0 ** 2)


KATS_IN = np.array([[1.0, 22.0, 0.0], [12.0, 23.0, 0.0], [10.0, 8.0, 2.0], [10.0, 0.0, 1.0]])
KATS_OUT = np.array([[12185139.590523569, 7420953.944297638, 0.0], [11104667.534080556, 7617693.973680517, 0.0],
                     [11055663.927418157, 5832081.512011217, 2.0], [10807262.110686881, 4909128.916889962, 1.0]])


@pytest.mark.parametrize("columnar,device", [(False, "cuda"), (True, "cuda"), (False, "cpu")])
def test_reproject_epsg4326_epsg3309_within(columnar, device):  # reprojection.rs:250-291
    buf = cloud(KATS_IN, columnar, device, extra=True)
    reproject_point_cloud_within(buf, "EPSG:4326", "EPSG:3309")
    torch.cuda.synchronize()
    out = buf.view_attribute(A.POSITION_3D)
    assert np.all(np.abs(out - KATS_OUT) < 1e-4), out - KATS_OUT
    assert np.array_equal(buf.view_attribute(A.INTENSITY), np.zeros(4, np.uint16))


def test_reproject_between_and_errors():  # reprojection.rs:292-367
    src = cloud(KATS_IN, False, extra=True)
    dst = cloud(np.zeros((4, 3)), True, extra=True)
    reproject_point_cloud_between(src, dst, "EPSG:4326", "EPSG:3309")
    torch.cuda.synchronize()
    assert np.all(np.abs(dst.view_attribute(A.POSITION_3D) - KATS_OUT) < 1e-4)
    assert np.array_equal(src.view_attribute(A.POSITION_3D), KATS_IN)
    with pytest.raises(pb.PastureB200Error) as e:  # "The point clouds don't have the same size!"
        reproject_point_cloud_between(src, cloud(np.zeros((2, 3))), "EPSG:4326", "EPSG:3309")
    assert e.value.code == -5
    with pytest.raises(pb.PastureB200Error) as e:
        reproject_point_cloud_within(src, "EPSG:4326", "EPSG:27700")  # needs a datum grid: no built-in pipeline
    assert e.value.code == -10


def test_reproject_matches_oracle_on_random_points():
    rng = np.random.default_rng(5)
    pts = np.stack([rng.random(20000) * 60 - 10, rng.random(20000) * 80 - 160, rng.random(20000) * 3000], 1)
    ops, n = O.pipeline_epsg4326_to_3309()
    expect = O.reproject(ops, n, pts)
    buf = cloud(pts)
    reproject_point_cloud_within(buf, "EPSG:4326", "EPSG:3309")
    torch.cuda.synchronize()
    got = buf.view_attribute(A.POSITION_3D)
    assert np.all(np.abs(got - expect) < 1e-6), np.abs(got - expect).max()
    assert np.array_equal(got[:, 2], pts[:, 2])  # z passes through


@pytest.mark.parametrize("target,zone,south,ell", [("EPSG:32632", 32, False, O.WGS84), ("EPSG:32718", 18, True, O.WGS84),
                                                   ("EPSG:25832", 32, False, O.GRS80), ("EPSG:32601", 1, False, O.WGS84)])
def test_reproject_utm_zones_match_oracle_both_ways(target, zone, south, ell):
    """EPSG:4326 <-> UTM (WGS 84 north / south, ETRS89): the GPU pipeline against the oracle's restatement of the Guidance
    Note 7-2 formulas (pinned on the published worked examples in the CPU suite); points over the whole zone width"""
    rng = np.random.default_rng(zone)
    n = 50_000
    lat = rng.uniform(-79.0, -1.0, n) if south else rng.uniform(1.0, 83.0, n)
    pts = np.stack([lat, 6.0 * zone - 183.0 + rng.uniform(-3.5, 3.5, n), rng.uniform(-50, 4000, n)], 1)
    fwd, nf = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.utm_step(zone, south, ell)])
    expect = O.reproject(fwd, nf, pts)
    buf = cloud(pts)
    reproject_point_cloud_within(buf, "EPSG:4326", target)
    torch.cuda.synchronize()
    got = buf.view_attribute(A.POSITION_3D)
    assert np.max(np.abs(got - expect)) < 1e-6, np.max(np.abs(got - expect))  # device libm vs glibc: nanometres
    assert np.array_equal(got[:, 2], pts[:, 2])
    reproject_point_cloud_within(buf, target, "EPSG:4326")  # and back: the reverse series
    torch.cuda.synchronize()
    back = buf.view_attribute(A.POSITION_3D)
    assert np.max(np.abs(back[:, :2] - pts[:, :2])) < 1e-9
    inv, ni = O.make_pipeline([O.utm_step(zone, south, ell, inverse=True), (O.PROJ_RAD2DEG_LATLON, [])])
    assert np.max(np.abs(back - O.reproject(inv, ni, expect))) < 1e-9


def test_reproject_published_worked_examples_on_the_gpu():
    """the worked examples that pin the oracle (IOGP Guidance Note 7-2, Snyder), straight through the GPU kernel"""
    import ctypes as C
    from pasture_b200._lib import ProjOp, check, lib
    ctx = pb.get_context()

    def run(ops, pts):
        arr = (ProjOp * len(ops))(*ops)
        buf = cloud(np.asarray(pts, dtype=np.float64))
        d = buf.desc()
        check(lib().pb200_reproject(ctx._h, C.byref(d), None, arr, len(ops)))
        torch.cuda.synchronize()
        return buf.view_attribute(A.POSITION_3D)

    def tm(a, invf, lat0, lon0, k0, fe, fn, inverse=0):
        op = ProjOp()
        check(lib().pb200_proj_op_tmerc(a, invf, lat0, lon0, k0, fe, fn, inverse, C.byref(op)))
        return op
    d2r, r2d = ProjOp(), ProjOp()
    d2r.kind, r2d.kind = 8, 9
    osgb = (6377563.396, 299.32496, 49.0, -2.0, 0.9996012717, 400000.0, -100000.0)
    out = run([d2r, tm(*osgb)], [[50.5, 0.5, 0.0]])  # GN7-2 3.5.3.1
    assert abs(out[0, 0] - 577274.99) < 0.011 and abs(out[0, 1] - 69740.50) < 0.011
    back = run([tm(*osgb, inverse=1), r2d], out)
    assert abs(back[0, 0] - 50.5) < 1e-9 and abs(back[0, 1] - 0.5) < 1e-9
    out = run([d2r, tm(6378206.4, 294.978698213898, 0.0, -75.0, 0.9996, 500000.0, 0.0)], [[40.5, -73.5, 0.0]])  # Snyder, UTM 18
    assert abs(out[0, 0] - 627106.5) < 0.06 and abs(out[0, 1] - 4484124.4) < 0.06
    buf = cloud(np.array([[24 + 22 / 60 + 54.433 / 3600, -(100 + 20 / 60), 1.0]]))  # GN7-2 3.5.1.2
    reproject_point_cloud_within(buf, "EPSG:4326", "EPSG:3857")
    wm = buf.view_attribute(A.POSITION_3D)
    assert abs(wm[0, 0] + 11169055.58) < 0.01 and abs(wm[0, 1] - 2800000.00) < 0.01
    reproject_point_cloud_within(buf, "EPSG:3857", "EPSG:4326")
    ll = buf.view_attribute(A.POSITION_3D)
    assert abs(ll[0, 0] - (24 + 22 / 60 + 54.433 / 3600)) < 1e-10 and abs(ll[0, 1] + (100 + 20 / 60)) < 1e-10
    h = ProjOp()
    check(lib().pb200_proj_op_helmert(0.0, 0.0, 4.5, 0.0, 0.0, 0.554, 0.219, 0, C.byref(h)))  # GN7-2 4.3.3, WGS 72 -> WGS 84
    out = run([h], [[3657660.66, 255768.55, 5201382.11]])
    assert np.all(np.abs(out[0] - [3657660.78, 255778.43, 5201387.75]) < 0.01)
    cf = ProjOp()
    check(lib().pb200_proj_op_helmert(0.0, 0.0, 4.5, 0.0, 0.0, -0.554, 0.219, 1, C.byref(cf)))  # same in the Coordinate Frame convention
    assert np.allclose(run([cf], [[3657660.66, 255768.55, 5201382.11]]), out, atol=1e-9)


def test_reproject_utm_zone_to_zone():
    """points near a zone border given in zone 32 -> zone 33 (reverse + forward series): equals the direct projection"""
    rng = np.random.default_rng(9)
    pts = np.stack([rng.uniform(45, 55, 5000), rng.uniform(11.0, 13.0, 5000), np.zeros(5000)], 1)
    b32, b33 = cloud(pts), cloud(pts)
    reproject_point_cloud_within(b32, "EPSG:4326", "EPSG:32632")
    reproject_point_cloud_within(b33, "EPSG:4326", "EPSG:32633")
    reproject_point_cloud_within(b32, "EPSG:32632", "EPSG:32633")
    torch.cuda.synchronize()
    assert np.max(np.abs(b32.view_attribute(A.POSITION_3D) - b33.view_attribute(A.POSITION_3D))) < 1e-6


def test_radius_search_large_cloud_prunes_by_radius():
    """400 k points, small radius, list never full: the traversal must prune by r^2 (not only by the k-th distance),
    and the result must match a grid-hash recomputation"""
    rng = np.random.default_rng(5)
    n, r, m = 400_000, 0.004, 16
    pts = rng.random((n, 3))
    idx, cnt = radius_search(cloud(pts), r, m)
    idx = idx.cpu().numpy().view(np.uint32)
    cnt = cnt.cpu().numpy()
    cell = np.floor(pts / r).astype(np.int64)
    order = np.lexsort((cell[:, 2], cell[:, 1], cell[:, 0]))
    keys = (cell[:, 0] << 40) | (cell[:, 1] << 20) | cell[:, 2]
    skeys = keys[order]
    for i in range(0, n, 9973):
        cand = []
        c = cell[i]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    kk = ((c[0] + dx) << 40) | ((c[1] + dy) << 20) | (c[2] + dz)
                    lo, hi = np.searchsorted(skeys, kk, "left"), np.searchsorted(skeys, kk, "right")
                    cand.extend(order[lo:hi].tolist())
        cand = np.array(sorted(cand))
        d = pts[cand] - pts[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        inside = cand[d2 <= r * r]
        d2i = d2[d2 <= r * r]
        expect = inside[np.lexsort((inside, d2i))][:m]
        assert cnt[i] == len(expect)
        assert np.array_equal(idx[i, : cnt[i]], expect)
