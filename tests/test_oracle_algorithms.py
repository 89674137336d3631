"""Pins the oracle's algorithm restatements (bounds, minmax, voxel grid, normals, reprojection, kNN)
against the reference's own tests."""
import numpy as np
import pytest

import oracle as O

POSITION_3D = ("Position3D", O.VEC3F64)


def positions_buffer(xyz, columnar=True, dtype=O.VEC3F64):
    l = O.OLayout.from_attributes([("Position3D", dtype)])
    b = O.OBuffer(l, len(xyz), columnar)
    if len(xyz):
        b.set_attribute("Position3D", xyz)
    return b


def test_aabb_from_iter():  # pasture-core/src/math/bounds.rs:304-315
    b = positions_buffer(np.array([[0, 0, 0], [1, 1, 1], [-1, -1, -1]], dtype=np.float64))
    mn, mx = O.calculate_bounds(b)
    assert list(mn) == [-1, -1, -1] and list(mx) == [1, 1, 1]


def test_bounds_none_cases():  # bounds.rs:12-21
    assert O.calculate_bounds(positions_buffer(np.zeros((0, 3)))) is None
    l = O.OLayout.from_attributes([("Intensity", O.U16)])
    assert O.calculate_bounds(O.OBuffer(l, 3, True)) is None


def test_bounds_custom_position_type_and_nan():  # bounds.rs:56-85, :34-51
    b = positions_buffer(np.array([[1, 2, 3], [-4, 5, 6]], dtype=np.int32), dtype=O.VEC3I32)
    mn, mx = O.calculate_bounds(b)
    assert list(mn) == [-4, 2, 3] and list(mx) == [1, 5, 6]
    b = positions_buffer(np.array([[1, np.nan, 3], [0, 5, np.nan], [2, 4, 1]]))
    mn, mx = O.calculate_bounds(b)
    assert list(mn) == [0, 4, 1] and list(mx) == [2, 5, 3]


def test_las_fixture_bounds(las_fixtures):  # pasture-io/src/las/test_util.rs:46-48
    entry = las_fixtures["plain"]["0"]
    raw = O.OLayout.las_raw(0)
    src = O.OBuffer(raw, 10, False)
    src.aos[:] = np.frombuffer(bytes.fromhex(entry["records_hex"]), dtype=np.uint8)
    target = O.OLayout.las_default(0)
    dst = O.OConverter.las_default(raw, target, entry["scale"], entry["offset"]).convert(src, True)
    mn, mx = O.calculate_bounds(dst)
    assert list(mn) == [0, 0, 0] and list(mx) == [9, 9, 9]


def test_minmax_attribute():  # minmax.rs:13-51, math/minmax.rs doc-tests :17-18,:29-30
    l = O.OLayout.from_attributes([POSITION_3D, ("Intensity", O.I16), ("GpsTime", O.F64)])
    b = O.OBuffer(l, 4, False)
    b.set_attribute("Position3D", [[1, 2, 3], [2, 1, 0], [0, 0, 9], [5, -5, 5]])
    b.set_attribute("Intensity", [5, -3, 7, 0])
    b.set_attribute("GpsTime", [1.5, -2.5, 0.0, 9.0])
    mn, mx = O.minmax_attribute(b, "Position3D", O.VEC3F64)
    assert list(mn) == [0, -5, 0] and list(mx) == [5, 2, 9]
    mn, mx = O.minmax_attribute(b, "Intensity", O.I16)
    assert (mn[0], mx[0]) == (-3, 7)
    mn, mx = O.minmax_attribute(b, "GpsTime", O.F64)
    assert (mn[0], mx[0]) == (-2.5, 9.0)
    with pytest.raises(O.OracleError):
        O.minmax_attribute(b, "Nope", O.U8)
    with pytest.raises(O.OracleError):  # T != attribute datatype always panics (buffer_views.rs:549)
        O.minmax_attribute(b, "Intensity", O.I16, O.I32)
    assert O.minmax_attribute(O.OBuffer(l, 0, False), "Intensity", O.I16) is None


# ---------------------------------------------------------------- voxel grid ---------------------------

COMPLETE = [POSITION_3D, ("Intensity", O.U16), ("ReturnNumber", O.U8), ("NumberOfReturns", O.U8),
            ("ClassificationFlags", O.U8), ("ScannerChannel", O.U8), ("ScanDirectionFlag", O.U8),
            ("EdgeOfFlightLine", O.U8), ("Classification", O.U8), ("ScanAngleRank", O.I8), ("ScanAngle", O.I16),
            ("UserData", O.U8), ("PointSourceID", O.U16), ("ColorRGB", O.VEC3U16), ("GpsTime", O.F64),
            ("NIR", O.U16)]


def setup_point_cloud(seed=0):
    """voxel_grid.rs:757-904 (random fields drawn from the same ranges)"""
    rng = np.random.default_rng(seed)
    n = 3002
    pos = np.zeros((n, 3))
    pos[0] = 0.0
    pos[1] = 10.0
    intensity = rng.integers(200, 800, n)
    retn = rng.integers(20, 80, n)
    cflags = rng.integers(7, 20, n)
    sdf = rng.integers(0, 47, n)
    idx = 2
    for i in range(10):
        for j in range(10):
            for k in range(10):
                for d, inten, rn, cf, sd in ((0.5, 2, 32, 3, 0), (0.6, 4, 42, 7, 0), (0.7, 6, 42, 133, 1)):
                    pos[idx] = (i + d, j + d, k + d)
                    intensity[idx], retn[idx], cflags[idx], sdf[idx] = inten, rn, cf, sd
                    idx += 1
    l = O.OLayout.from_attributes(COMPLETE, packed=1)
    b = O.OBuffer(l, n, True)
    b.set_attribute("Position3D", pos)
    b.set_attribute("Intensity", intensity)
    b.set_attribute("ReturnNumber", retn)
    b.set_attribute("NumberOfReturns", rng.integers(20, 80, n))
    b.set_attribute("ClassificationFlags", cflags)
    b.set_attribute("ScannerChannel", rng.integers(7, 20, n))
    b.set_attribute("ScanDirectionFlag", sdf)
    b.set_attribute("EdgeOfFlightLine", rng.integers(0, 81, n))
    b.set_attribute("Classification", rng.integers(121, 200, n))
    b.set_attribute("ScanAngleRank", rng.integers(-121, 20, n))
    b.set_attribute("ScanAngle", rng.integers(-21, 8, n))
    b.set_attribute("UserData", rng.integers(1, 8, n))
    b.set_attribute("PointSourceID", rng.integers(9, 89, n))
    col = rng.integers(11, 120, (n, 3))
    col[:, 2] = 42
    b.set_attribute("ColorRGB", col)
    b.set_attribute("GpsTime", rng.random(n) * 103.7 - 22.4)
    b.set_attribute("NIR", rng.integers(4, 82, n))
    return l, b


@pytest.mark.parametrize("use_sort", [False, True])
def test_voxel_grid_filter(use_sort):  # voxel_grid.rs:906-938
    l, buf = setup_point_cloud()
    out, keys = O.voxelgrid_filter(buf, (1.0, 1.0, 1.0), l, columnar=True, use_sort=use_sort)
    assert out.len == 1000
    p = out.attribute("Position3D")[1]
    assert 0.59 < p[0] < 0.61 and 0.59 < p[1] < 0.61 and 1.59 < p[2] < 1.61
    assert out.attribute("Intensity")[1] == 4
    assert out.attribute("ReturnNumber")[1] == 42
    assert out.attribute("ClassificationFlags")[1] == 133
    assert tuple(keys[1]) == (0, 0, 1)
    assert out.attribute("ScanDirectionFlag")[1] == 0  # mode of (0,0,1) != 0 -> false
    assert out.attribute("ColorRGB")[1][2] == 42
    # lexicographic (ix,iy,iz) order, strictly increasing
    kk = [tuple(int(x) for x in k) for k in keys]
    assert kk == sorted(kk) and len(set(kk)) == len(kk)


def test_voxel_sort_variant_equals_faithful():
    rng = np.random.default_rng(5)
    n = 5000
    l = O.OLayout.from_attributes([POSITION_3D, ("Intensity", O.U16), ("Classification", O.U8),
                                   ("GpsTime", O.F64), ("ColorRGB", O.VEC3U16)])
    b = O.OBuffer(l, n, True)
    b.set_attribute("Position3D", rng.random((n, 3)) * [20, 10, 3])
    b.set_attribute("Intensity", rng.integers(0, 65536, n))
    b.set_attribute("Classification", rng.integers(0, 6, n))
    b.set_attribute("GpsTime", rng.random(n) * 100 - 20)
    b.set_attribute("ColorRGB", rng.integers(0, 65536, (n, 3)))
    a, ka = O.voxelgrid_filter(b, (0.7, 0.9, 1.1), l, True, use_sort=False)
    c, kc = O.voxelgrid_filter(b, (0.7, 0.9, 1.1), l, False, use_sort=True)
    assert a.len == c.len and np.array_equal(ka, kc)
    for i in range(l.n):
        assert np.array_equal(a.attribute_bytes(i), c.attribute_bytes(i))


def test_find_leaf_nearest_marker_rule():
    """voxel_grid.rs:22-51 with leaf 1, min 0: nearest-marker fix-up, not floor"""
    m = O.create_markers(0.0, 5.0, 1.0)
    assert list(m) == [1, 2, 3, 4, 5]
    for x, expect in ((1.2, 0), (1.49, 0), (1.5, 1), (0.2, 0), (4.9, 4), (5.0, 4), (2.0, 1)):
        p = np.array([x, 0.0, 0.0])
        lin = O.find_leaf(p, m, m, m)
        assert lin[0] == expect, (x, lin)
        assert lin == O.find_leaf(p, m, m, m, bsearch=True)
    # cumulative-sum markers (0.1 steps accumulate rounding), linear == binary search on random points
    mk = O.create_markers(-3.0, 7.3, 0.1)
    assert mk[0] == -3.0 + 0.1 and mk[1] == (-3.0 + 0.1) + 0.1
    rng = np.random.default_rng(0)
    for x in rng.random(500) * 10.3 - 3.0:
        p = np.array([x, x, x])
        assert O.find_leaf(p, mk, mk, mk) == O.find_leaf(p, mk, mk, mk, bsearch=True)
    empty = np.zeros(0)
    assert O.find_leaf(np.array([1.0, 2.0, 3.0]), empty, empty, empty) == (0, 0, 0)


def test_voxel_unsupported_attributes_panic():  # voxel_grid.rs:452-459, :682-687
    src_l = O.OLayout.from_attributes([POSITION_3D, ("WaveformPacketSize", O.U32), ("Custom", O.F32)])
    b = O.OBuffer(src_l, 3, True)
    b.set_attribute("Position3D", [[0, 0, 0], [1, 1, 1], [2, 2, 2]])
    with pytest.raises(O.OracleError):
        O.voxelgrid_filter(b, (1, 1, 1), O.OLayout.from_attributes([POSITION_3D, ("WaveformPacketSize", O.U32)]))
    with pytest.raises(O.OracleError):
        O.voxelgrid_filter(b, (1, 1, 1), O.OLayout.from_attributes([POSITION_3D, ("Custom", O.F32)]))


# ---------------------------------------------------------------- normals -------------------------------

KAT = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])


def test_compute_normal_sub():  # normal_estimation.rs:503-550
    assert list(O.compute_centroid(KAT)) == [0.25, 0.5, 0.0]
    cov = O.compute_covariance(KAT)
    assert np.array_equal(cov, np.array([[2.75, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.0]]))
    n, curv = O.solve_plane_parameter(cov)
    assert n[0] == 0.0 and n[1] == 0.0 and n[2] != 0.0 and curv == 0.0


def test_covariance_error():  # :552-578
    nan = np.nan
    pts = np.array([[nan, 0, 0], [0, 1, nan], [1, 1, nan], [-1, nan, 0]])
    with pytest.raises(O.OracleError) as e:
        O.compute_covariance(pts)
    assert e.value.code == O.ERR_TOO_FEW_POINTS


def test_compute_normal():  # :580-610
    normals, curv = O.compute_normals(KAT, 3)
    assert np.all(normals[:, 0] == 0) and np.all(normals[:, 1] == 0) and np.all(normals[:, 2] != 0)
    assert np.all(curv == 0)


def test_compute_normal_panics():  # :612-698
    with pytest.raises(O.OracleError) as e:
        O.compute_normals(KAT[:1], 3)
    assert e.value.code == O.ERR_TOO_FEW_POINTS
    with pytest.raises(O.OracleError):
        O.compute_normals(KAT[:2], 3)
    for k in (1, 2):
        with pytest.raises(O.OracleError) as e:
            O.compute_normals(KAT, k)
        assert e.value.code == O.ERR_INVALID


def test_normals_follow_the_code_as_written():
    """normal = largest cross product of rows of C/scale (no eigenvalue shift);
    curvature = |lambda0(C) * maxabs(C) / trace(C)|"""
    rng = np.random.default_rng(3)
    pts = rng.random((16, 3)) * [1, 1, 0.05]
    cov = O.compute_covariance(pts)
    n, curv = O.solve_plane_parameter(cov)
    S = cov / np.abs(cov).max()
    rows = [np.cross(S[0], S[1]), np.cross(S[0], S[2]), np.cross(S[1], S[2])]
    best = rows[0]
    for r in rows:
        if np.linalg.norm(r) > np.linalg.norm(best):
            best = r
    assert np.allclose(n, best, rtol=1e-12, atol=0)
    lam0 = np.linalg.eigvalsh(cov)[0]
    assert curv == pytest.approx(abs(lam0 * np.abs(cov).max() / np.trace(cov)), rel=1e-6)
    assert abs(np.dot(n / np.linalg.norm(n), [0, 0, 1])) > 0.9


def test_knn_bruteforce_includes_query_and_is_sorted():
    rng = np.random.default_rng(2)
    pts = rng.random((300, 3))
    idx, d2 = O.knn_bruteforce(pts, pts[:20], 8)
    assert np.array_equal(idx[:, 0], np.arange(20)) and np.all(d2[:, 0] == 0)
    assert np.all(np.diff(d2, axis=1) >= 0)
    ref = np.argsort(((pts[None, :, :] - pts[:20, None, :]) ** 2).sum(-1), axis=1, kind="stable")[:, :8]
    assert np.array_equal(np.sort(idx, axis=1), np.sort(ref, axis=1))


# ---------------------------------------------------------------- reprojection ------------------------

def test_reproject_epsg4326_epsg3309():  # reprojection.rs:250-337 (assert_approx_eq 1e-4)
    pts = np.array([[1.0, 22.0, 0.0], [12.0, 23.0, 0.0], [10.0, 8.0, 2.0], [10.0, 0.0, 1.0]])
    expected = np.array([[12185139.590523569, 7420953.944297638, 0.0],
                         [11104667.534080556, 7617693.973680517, 0.0],
                         [11055663.927418157, 5832081.512011217, 2.0],
                         [10807262.110686881, 4909128.916889962, 1.0]])
    ops, n = O.pipeline_epsg4326_to_3309()
    out = O.reproject(ops, n, pts)
    assert np.all(np.abs(out - expected) < 1e-4), out - expected


# Published worked examples stand in for libproj (proj-sys 0.22 is not in /root/reference): the Transverse
# Mercator, Pseudo-Mercator and Helmert operations are pinned on IOGP Guidance Note 7-2 and Snyder (USGS PP 1395).

def _dms(d, m, s):
    return d + m / 60.0 + s / 3600.0


def test_transverse_mercator_guidance_note_example():
    """GN7-2 3.5.3.1, OSGB 1936 / British National Grid: Airy 1830 (a = 6377563.396, 1/f = 299.32496), origin 49 N 2 W,
    k0 = 0.9996012717, FE 400000, FN -100000; 50 30' N 0 30' E -> E 577274.99, N 69740.50 (published to the centimetre)"""
    airy = (6377563.396, 299.32496)
    ops, n = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.tmerc_step(airy, 49.0, -2.0, 0.9996012717, 400000.0, -100000.0)])
    out = O.reproject(ops, n, np.array([[50.5, 0.5, 7.0]]))
    assert abs(out[0, 0] - 577274.99) < 0.011 and abs(out[0, 1] - 69740.50) < 0.011 and out[0, 2] == 7.0
    ops, n = O.make_pipeline([O.tmerc_step(airy, 49.0, -2.0, 0.9996012717, 400000.0, -100000.0, inverse=True), (O.PROJ_RAD2DEG_LATLON, [])])
    back = O.reproject(ops, n, out)  # the note's reverse example returns to 50 30' 00.000" N, 00 30' 00.000" E
    assert abs(back[0, 0] - 50.5) < 1e-9 and abs(back[0, 1] - 0.5) < 1e-9


def test_utm_snyder_example_and_meridian_arc():
    """Snyder, Map Projections - A Working Manual, numerical example for the ellipsoidal Transverse Mercator: Clarke 1866,
    40 30' N 73 30' W in UTM zone 18 -> x = 127106.5 m east of the central meridian, y = 4484124.4 m.  Plus the meridian
    arc (the series at lon = lon0) against numerical quadrature of the arc-length integral on WGS 84."""
    from scipy.integrate import quad
    clarke = (6378206.4, 294.978698213898)
    ops, n = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.utm_step(18, ellipsoid=clarke)])
    out = O.reproject(ops, n, np.array([[40.5, -73.5, 0.0]]))
    assert abs(out[0, 0] - 627106.5) < 0.06 and abs(out[0, 1] - 4484124.4) < 0.06
    a, invf = O.WGS84
    f = 1.0 / invf
    e2 = f * (2.0 - f)
    ops, n = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.tmerc_step(O.WGS84, 0.0, 0.0, 1.0, 0.0, 0.0)])
    for lat in (5.0, 33.0, 60.0, 84.0):
        arc = quad(lambda p: a * (1.0 - e2) / (1.0 - e2 * np.sin(p) ** 2) ** 1.5, 0.0, np.radians(lat), epsabs=1e-9)[0]
        out = O.reproject(ops, n, np.array([[lat, 0.0, 0.0]]))
        assert abs(out[0, 1] - arc) < 1e-6 and abs(out[0, 0]) < 1e-9


def test_utm_round_trip_and_zone_parameters():
    """forward and reverse series have independent coefficients: a round trip over a whole zone (+-4 deg) to 1e-10 deg;
    zone 32 north has its central meridian at 9 E, false easting 500 km; the southern hemisphere adds 10 000 km"""
    rng = np.random.default_rng(7)
    pts = np.stack([rng.uniform(-80, 84, 3000), 9.0 + rng.uniform(-4, 4, 3000), rng.uniform(0, 500, 3000)], axis=1)
    fwd, nf = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.utm_step(32)])
    inv, ni = O.make_pipeline([O.utm_step(32, inverse=True), (O.PROJ_RAD2DEG_LATLON, [])])
    en = O.reproject(fwd, nf, pts)
    back = O.reproject(inv, ni, en)
    assert np.max(np.abs(back - pts)) < 1e-10
    on_cm = O.reproject(fwd, nf, np.array([[0.0, 9.0, 0.0], [45.0, 9.0, 0.0]]))
    assert abs(on_cm[0, 0] - 500000.0) < 1e-9 and abs(on_cm[0, 1]) < 1e-9 and abs(on_cm[1, 0] - 500000.0) < 1e-9
    south, ns = O.make_pipeline([(O.PROJ_DEG2RAD_LATLON, []), O.utm_step(32, south=True)])
    assert abs(O.reproject(south, ns, np.array([[0.0, 9.0, 0.0]]))[0, 1] - 10000000.0) < 1e-9


def test_pseudo_mercator_guidance_note_example():
    """GN7-2 3.5.1.2 (EPSG method 1024, WGS 84 / Pseudo-Mercator): 24 22' 54.433" N 100 20' W -> E -11169055.58, N 2800000.00"""
    ops, n = O.make_pipeline([(O.PROJ_WEBMERC_FWD, [])])
    out = O.reproject(ops, n, np.array([[_dms(24, 22, 54.433), -_dms(100, 20, 0.0), 3.0]]))
    assert abs(out[0, 0] + 11169055.58) < 0.01 and abs(out[0, 1] - 2800000.00) < 0.01 and out[0, 2] == 3.0
    inv, ni = O.make_pipeline([(O.PROJ_WEBMERC_INV, [])])
    back = O.reproject(inv, ni, out)
    assert abs(back[0, 0] - _dms(24, 22, 54.433)) < 1e-10 and abs(back[0, 1] + _dms(100, 20, 0.0)) < 1e-10


def test_helmert_guidance_note_example():
    """GN7-2 4.3.3 Position Vector transformation WGS 72 -> WGS 84: tZ = +4.5 m, rZ = +0.554", dS = +0.219 ppm;
    (3657660.66, 255768.55, 5201382.11) -> (3657660.78, 255778.43, 5201387.75)"""
    ops, n = O.make_pipeline([O.helmert_step(0.0, 0.0, 4.5, 0.0, 0.0, 0.554, 0.219)])
    out = O.reproject(ops, n, np.array([[3657660.66, 255768.55, 5201382.11]]))
    assert np.all(np.abs(out[0] - [3657660.78, 255778.43, 5201387.75]) < 0.01), out


# ---- the oracle's two kNN implementations pin each other -------------------------------------------------------------

@pytest.mark.parametrize("seed,n,k", [(0, 5000, 16), (1, 777, 3), (2, 40, 64), (3, 9000, 33)])
def test_kdtree_knn_equals_brute_force(seed, n, k):
    """po_knn_kdtree (exact kd-tree, the checker for clouds the O(N^2) brute force cannot reach) returns the same indices
    and bit-identical distances as po_knn_bruteforce, duplicates and exact ties included ((d2, index) order)"""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 3)) * [30.0, 30.0, 2.0]
    pts[n // 3] = pts[1]
    pts[n // 2: n // 2 + 7] = pts[2]                      # a cluster of identical points
    pts[5:25, 2] = 0.5                                    # a plane of equal z: ties in one coordinate
    bi, bd = O.knn_bruteforce(pts, pts, k)
    ki, kd = O.knn_kdtree(pts, k, threads=3)
    assert np.array_equal(bi, ki) and np.array_equal(bd, kd)
    lo, hi = n // 4, min(n, n // 4 + 50)
    ri, rd = O.knn_kdtree(pts, k, lo, hi, threads=2)      # query range: rows of the full result
    assert np.array_equal(ri, bi[lo:hi]) and np.array_equal(rd, bd[lo:hi])


def test_kdtree_normals_equal_brute_force_normals():
    pts = O.gen_terrain_positions(0, 4000)
    n1, c1 = O.compute_normals(pts, 16)
    n2, c2 = O.compute_normals_kdtree(pts, 16)
    assert np.array_equal(n1, n2) and np.array_equal(c1, c2)
