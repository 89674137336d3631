"""Per-point algorithm sweeps -- host-side mirror of pasture-algorithms (bounds.rs, minmax.rs, voxel_grid.rs,
normal_estimation.rs, reprojection.rs) and pasture-core/src/math (bounds.rs AABB, bitmanip.rs)."""
import ctypes as C

import numpy as np
import torch

from ._lib import BufferDesc, ProjOp, VoxelAttrPartialsDesc, VoxelPartialsDesc, check, lib
from .containers import _NP, _VEC3, HashMapBuffer, VectorBuffer
from .context import context_for, get_context
from .layout import DT


class AABB:
    """math/bounds.rs:9"""

    def __init__(self, mn, mx):
        self._min, self._max = tuple(mn), tuple(mx)

    def min(self):
        return self._min

    def max(self):
        return self._max

    def extent(self):
        return tuple(b - a for a, b in zip(self._min, self._max))

    @staticmethod
    def union(a, b):
        return AABB([min(x, y) for x, y in zip(a._min, b._min)], [max(x, y) for x, y in zip(a._max, b._max)])

    def __eq__(self, o):
        return isinstance(o, AABB) and self._min == o._min and self._max == o._max

    def __repr__(self):
        return f"AABB(min={self._min}, max={self._max})"


def calculate_bounds(buffer, ctx=None):
    """bounds.rs:11-28 -> AABB or None"""
    ctx = context_for(ctx, buffer)
    d = buffer.desc()
    mn, mx, some = (C.c_double * 3)(), (C.c_double * 3)(), C.c_int(0)
    check(lib().pb200_calculate_bounds(ctx._h, C.byref(d), mn, mx, C.byref(some)))
    return AABB(mn, mx) if some.value else None


def minmax_attribute(buffer, attribute, ctx=None, partial=False):
    """minmax.rs:13-51 -> (min, max) of attribute.datatype() or None.  partial=True: the continuation fold for a shard that
    does not start the cloud (NaN never enters; see pb200_minmax_attribute_partial and sharding.minmax_attribute_sharded)"""
    ctx = context_for(ctx, buffer)
    d = buffer.desc()
    mn, mx, some = np.zeros(32, np.uint8), np.zeros(32, np.uint8), C.c_int(0)
    fn = lib().pb200_minmax_attribute_partial if partial else lib().pb200_minmax_attribute
    check(fn(ctx._h, C.byref(d), attribute.name().encode(), int(attribute.datatype()),
             C.c_void_p(mn.ctypes.data), C.c_void_p(mx.ctypes.data), C.byref(some)))
    if not some.value:
        return None
    dt = attribute.datatype()
    comp = _NP.get(dt) or _NP[_VEC3[dt]]
    sz = attribute.size()
    a, b = mn[:sz].view(comp).copy(), mx[:sz].view(comp).copy()
    return (a[0], b[0]) if dt in _NP else (a, b)


def expand_bits_by_3(v):
    """math/bitmanip.rs:2-10"""
    return int(lib().pb200_expand_bits_by_3(int(v) & 0xFFFFFFFFFFFFFFFF))


def morton_codes(buffer, bmin, bmax, ctx=None):
    ctx = context_for(ctx, buffer)
    d = buffer.desc()
    out = torch.zeros(max(1, buffer.len()), dtype=torch.int64, device=buffer.device)
    check(lib().pb200_morton_codes(ctx._h, C.byref(d), (C.c_double * 3)(*bmin), (C.c_double * 3)(*bmax),
                                   C.c_void_p(out.data_ptr())))
    return out[: buffer.len()]


class _ResultOwner:
    """keeps a library-owned result (pb200_result_buffer) alive for as long as tensors view its device memory"""

    def __init__(self, handle, ctx):
        self._h = handle
        self._ctx = ctx  # the context must outlive the result (its memory goes back to the context's cache)

    def __del__(self):
        try:
            if self._h:
                lib().pb200_result_buffer_destroy(self._h)
                self._h = None
        except Exception:
            pass


class _DeviceView:
    """a span of library-owned device memory for torch.as_tensor (CUDA array interface); holds the owner"""

    def __init__(self, ptr, nbytes, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def voxelgrid_filter(buffer, leafsize_x, leafsize_y, leafsize_z, filtered_layout=None, out_buffer_type=HashMapBuffer,
                     device=None, ctx=None, return_keys=False):
    """voxel_grid.rs:109-165. Returns the filtered buffer.  A device-resident result is handed over WITHOUT a copy: the
    returned buffer's tensors view the memory of the library's result object, which lives as long as they do (then the
    block goes back to the context's cache).  Host results are copied into torch-owned memory."""
    ctx = context_for(ctx, buffer)
    layout = filtered_layout or buffer.point_layout()
    dev = torch.device(device) if device is not None else buffer.device
    kind = 0 if out_buffer_type is VectorBuffer else 1
    memspace = 1 if dev.type == "cuda" else 0
    d = buffer.desc()
    h = C.c_void_p()
    check(lib().pb200_voxelgrid_filter(ctx._h, C.byref(d), leafsize_x, leafsize_y, leafsize_z, layout._h, kind, memspace,
                                       C.byref(h)))
    owner = _ResultOwner(h, ctx)
    rd = BufferDesc()
    check(lib().pb200_result_buffer_desc(h, C.byref(rd)))
    n = int(rd.len)
    if memspace == 1 and n:
        cdev = torch.device("cuda", ctx.device)

        def view(ptr, nbytes):
            return torch.as_tensor(_DeviceView(ptr, max(1, nbytes), owner), device=cdev)
        if kind == 0:
            out = VectorBuffer(layout, n, cdev, data=view(rd.aos, n * layout.size_of_point_entry()))
        else:
            out = HashMapBuffer(layout, n, cdev, columns=[view(rd.columns[i], n * a.size()) for i, a in enumerate(layout.attributes())])
    else:
        out = out_buffer_type(layout, n, dev)
        od = out.desc()
        if n:
            if kind == 0:
                _copy(ctx, od.aos, rd.aos, n * layout.size_of_point_entry(), memspace)
            else:
                for i, a in enumerate(layout.attributes()):
                    _copy(ctx, od.columns[i], rd.columns[i], n * a.size(), memspace)
    keys = None
    if return_keys:
        keys = np.zeros((max(1, n), 3), dtype=np.uint64)
        check(lib().pb200_result_buffer_voxel_keys(h, C.c_void_p(keys.ctypes.data)))
        keys = keys[:n]
    return (out, keys) if return_keys else out


def radix_sort(keys, vals=None, begin_bit=0, end_bit=64, ctx=None):
    """K8: stable LSD radix sort of a CUDA int64 tensor (u64 bit patterns) by the bits [begin_bit, end_bit), in place,
    optionally carrying an int32 payload (stability: equal keys keep their input order)"""
    ctx = context_for(ctx, keys)
    assert keys.dtype == torch.int64 and keys.is_contiguous() and keys.is_cuda
    assert vals is None or (vals.dtype == torch.int32 and vals.is_contiguous() and vals.numel() == keys.numel())
    check(lib().pb200_radix_sort_u64(ctx._h, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()) if vals is not None else None,
                                     keys.numel(), begin_bit, end_bit))
    return keys if vals is None else (keys, vals)


class VoxelPartials:
    """Per-voxel partial sums on a (global) voxel grid as device tensors: packed keys (int64 bit pattern of the u64 key
    (ix << (bits_y + bits_z)) | (iy << bits_z) | iz, ascending), point counts (int32) and position sums [V, 3] f64.
    `bits` = (bits_x, bits_y, bits_z), `cells` = markers per axis.  SURVEY 8e."""

    def __init__(self, keys, counts, sums, bits, cells):
        self.keys, self.counts, self.sums, self.bits, self.cells = keys, counts, sums, tuple(bits), tuple(cells)

    def len(self):
        return int(self.keys.numel())

    def unpack_keys(self):
        """-> [V, 3] int64 voxel indices (ix, iy, iz)"""
        _, by, bz = self.bits
        k = self.keys
        return torch.stack([k >> (by + bz), (k >> bz) & ((1 << by) - 1), k & ((1 << bz) - 1)], 1)


def _take_partials(ctx, h, want_centroids):
    """copy a library-owned pb200_voxel_partials into torch tensors (stream-ordered) and release it"""
    try:
        d = VoxelPartialsDesc()
        check(lib().pb200_voxel_partials_get(h, C.byref(d)))
        v = int(d.len)
        dev = torch.device("cuda", ctx.device)
        keys = torch.empty(v, dtype=torch.int64, device=dev)
        counts = torch.empty(v, dtype=torch.int32, device=dev)
        sums = torch.empty((v, 3), dtype=torch.float64, device=dev)
        cent = torch.empty((v, 3), dtype=torch.float64, device=dev) if want_centroids else None
        if v:
            check(lib().pb200_memcpy_d2d(ctx._h, C.c_void_p(keys.data_ptr()), C.c_void_p(d.keys), 8 * v))
            check(lib().pb200_memcpy_d2d(ctx._h, C.c_void_p(counts.data_ptr()), C.c_void_p(d.counts), 4 * v))
            check(lib().pb200_memcpy_d2d(ctx._h, C.c_void_p(sums.data_ptr()), C.c_void_p(d.sums), 24 * v))
            if want_centroids:
                check(lib().pb200_voxel_partials_centroids(h, C.c_void_p(cent.data_ptr())))
        part = VoxelPartials(keys, counts, sums, (d.bits_x, d.bits_y, d.bits_z), tuple(d.cells))
    finally:
        lib().pb200_voxel_partials_destroy(h)
    return (part, cent) if want_centroids else part


def voxelgrid_partials(buffer, leafsize_x, leafsize_y, leafsize_z, global_bounds, ctx=None):
    """one shard's contribution to a sharded voxel grid: partial sums on the grid of `global_bounds` (an AABB or a
    (min, max) pair covering the WHOLE cloud, e.g. the all-reduced shard bounds)"""
    ctx = context_for(ctx, buffer)
    mn, mx = (global_bounds.min(), global_bounds.max()) if isinstance(global_bounds, AABB) else global_bounds
    gmin, gmax = (C.c_double * 3)(*mn), (C.c_double * 3)(*mx)
    d = buffer.desc()
    h = C.c_void_p()
    check(lib().pb200_voxelgrid_partials(ctx._h, C.byref(d), leafsize_x, leafsize_y, leafsize_z, gmin, gmax, C.byref(h)))
    return _take_partials(ctx, h, False)


class VoxelAttrPartials:
    """one shard's (or several concatenated shards') partials for every attribute of a filtered layout: position partials
    (VoxelPartials) + per-voxel columns [len, n_columns] + one (keys, counts) run list per "most common value" attribute"""

    def __init__(self, pos, columns, column_is_max, modes):
        self.pos, self.columns, self.column_is_max, self.modes = pos, columns, column_is_max, modes


def voxelgrid_partials_layout(buffer, leafsize_x, leafsize_y, leafsize_z, global_bounds, filtered_layout=None, ctx=None):
    """pb200_voxelgrid_partials_layout: everything one shard contributes to the filtered cloud in `filtered_layout`"""
    ctx = context_for(ctx, buffer)
    layout = filtered_layout or buffer.point_layout()
    mn, mx = (global_bounds.min(), global_bounds.max()) if isinstance(global_bounds, AABB) else global_bounds
    gmin, gmax = (C.c_double * 3)(*mn), (C.c_double * 3)(*mx)
    d = buffer.desc()
    h = C.c_void_p()
    check(lib().pb200_voxelgrid_partials_layout(ctx._h, C.byref(d), leafsize_x, leafsize_y, leafsize_z, gmin, gmax, layout._h, C.byref(h)))
    try:
        ad = VoxelAttrPartialsDesc()
        check(lib().pb200_voxel_partials_get_attrs(h, C.byref(ad)))
        pd = VoxelPartialsDesc()
        check(lib().pb200_voxel_partials_get(h, C.byref(pd)))
        v = int(pd.len)
        dev = torch.device("cuda", ctx.device)

        def take(ptr, shape, dtype):
            t = torch.empty(shape, dtype=dtype, device=dev)
            if t.numel():
                check(lib().pb200_memcpy_d2d(ctx._h, C.c_void_p(t.data_ptr()), C.c_void_p(ptr), t.numel() * t.element_size()))
            return t
        cols = take(ad.columns, (v, int(ad.n_columns)), torch.float64)
        modes = [(take(ad.mode_keys[a], (int(ad.mode_len[a]),), torch.int64), take(ad.mode_counts[a], (int(ad.mode_len[a]),), torch.int32))
                 for a in range(int(ad.n_modes))]
        pos = VoxelPartials(take(pd.keys, (v,), torch.int64), take(pd.counts, (v,), torch.int32), take(pd.sums, (v, 3), torch.float64),
                            (pd.bits_x, pd.bits_y, pd.bits_z), tuple(pd.cells))
        return VoxelAttrPartials(pos, cols, bytes(ad.column_is_max), modes)
    finally:
        lib().pb200_voxel_partials_destroy(h)


def voxelgrid_merge_partials_layout(partials, filtered_layout, ctx=None, return_keys=False):
    """merge concatenated VoxelAttrPartials (rows in source-rank order) -> the filtered points of this key range as a
    HashMapBuffer viewing the library-owned result (no copy)"""
    ctx = context_for(ctx, partials.pos.keys)
    pos = partials.pos
    keys, counts, sums = pos.keys.contiguous(), pos.counts.contiguous(), pos.sums.contiguous()
    cols = partials.columns.contiguous()
    pd = VoxelPartialsDesc()
    pd.len = keys.numel()
    pd.keys, pd.counts, pd.sums = keys.data_ptr(), counts.data_ptr(), sums.data_ptr()
    pd.bits_x, pd.bits_y, pd.bits_z = pos.bits
    ad = VoxelAttrPartialsDesc()
    ad.n_columns = cols.shape[1] if cols.dim() == 2 else 0
    ad.n_modes = len(partials.modes)
    ad.columns = cols.data_ptr() if cols.numel() else None
    for i, b in enumerate(partials.column_is_max[:64]):
        ad.column_is_max[i] = b
    keep = []
    for a, (mk, mc) in enumerate(partials.modes):
        mk, mc = mk.contiguous(), mc.contiguous()
        keep.append((mk, mc))
        ad.mode_len[a] = mk.numel()
        ad.mode_keys[a] = mk.data_ptr() if mk.numel() else None
        ad.mode_counts[a] = mc.data_ptr() if mc.numel() else None
    h = C.c_void_p()
    check(lib().pb200_voxelgrid_merge_partials_layout(ctx._h, filtered_layout._h, C.byref(pd), C.byref(ad), C.byref(h)))
    owner = _ResultOwner(h, ctx)
    rd = BufferDesc()
    check(lib().pb200_result_buffer_desc(h, C.byref(rd)))
    n = int(rd.len)
    cdev = torch.device("cuda", ctx.device)
    if n:
        out = HashMapBuffer(filtered_layout, n, cdev, columns=[torch.as_tensor(_DeviceView(rd.columns[i], max(1, n * a.size()), owner), device=cdev)
                                                                for i, a in enumerate(filtered_layout.attributes())])
    else:
        out = HashMapBuffer(filtered_layout, 0, cdev)
    if not return_keys:
        return out
    vk = np.zeros((max(1, n), 3), dtype=np.uint64)
    if n:
        check(lib().pb200_result_buffer_voxel_keys(h, C.c_void_p(vk.ctypes.data)))
    return out, vk[:n]


def voxelgrid_merge_partials(keys, counts, sums, bits, cells=(0, 0, 0), ctx=None):
    """merge concatenated partials (device tensors; equal keys are added in the order given = source-rank order)
    -> (VoxelPartials with the totals, centroids [V, 3])"""
    ctx = context_for(ctx, keys)
    keys, counts, sums = keys.contiguous(), counts.contiguous(), sums.contiguous()
    assert keys.dtype == torch.int64 and counts.dtype == torch.int32 and sums.dtype == torch.float64
    h = C.c_void_p()
    check(lib().pb200_voxelgrid_merge_partials(ctx._h, C.c_void_p(keys.data_ptr()), C.c_void_p(counts.data_ptr()),
                                               C.c_void_p(sums.data_ptr()), keys.numel(), bits[0], bits[1], bits[2], C.byref(h)))
    part, cent = _take_partials(ctx, h, True)
    part.cells = tuple(cells)
    return part, cent


def _copy(ctx, dst, src, nbytes, memspace):
    if nbytes == 0:
        return
    if memspace == 1:
        # the library-owned result buffer is released stream-ordered (behind this copy), so no synchronisation is needed
        check(lib().pb200_memcpy_d2d(ctx._h, C.c_void_p(dst), C.c_void_p(src), nbytes))
    else:
        C.memmove(dst, src, nbytes)


def knn(buffer, k, with_distances=True, query_range=None, ctx=None):
    """KdTree::nearests for every point (or the points of `query_range`) against the cloud (normal_estimation.rs:103,108)
    -> (idx [nq,k] int32 view of u32, d2 [nq,k]); row 0 belongs to query_range.start"""
    ctx = context_for(ctx, buffer)
    n = buffer.len()
    r = query_range if query_range is not None else range(0, n)
    nq = len(r)
    dev = buffer.device
    idx = torch.zeros((max(1, nq), k), dtype=torch.int32, device=dev)
    d2 = torch.zeros((max(1, nq), k), dtype=torch.float64, device=dev) if with_distances else None
    d = buffer.desc()
    check(lib().pb200_knn_range(ctx._h, C.byref(d), k, r.start, nq, C.c_void_p(idx.data_ptr()),
                                C.c_void_p(d2.data_ptr()) if with_distances else None))
    return (idx[:nq], d2[:nq]) if with_distances else idx[:nq]


def radius_search(buffer, radius, max_neighbors, ctx=None):
    ctx = context_for(ctx, buffer)
    n = buffer.len()
    dev = buffer.device
    idx = torch.zeros((max(1, n), max_neighbors), dtype=torch.int32, device=dev)
    cnt = torch.zeros(max(1, n), dtype=torch.int32, device=dev)
    d = buffer.desc()
    check(lib().pb200_radius_search(ctx._h, C.byref(d), radius, max_neighbors, C.c_void_p(idx.data_ptr()),
                                    C.c_void_p(cnt.data_ptr())))
    return idx[:n], cnt[:n]


def compute_normals(point_cloud, k_nn, query_range=None, ctx=None):
    """normal_estimation.rs:79-130 -> (normals [nq,3] f64, curvature [nq] f64) tensors on the buffer's device, for every
    point or for the points of `query_range` (the loop body of :106-127 over a sub-range; neighbours come from the whole
    cloud either way)"""
    ctx = context_for(ctx, point_cloud)
    n = point_cloud.len()
    r = query_range if query_range is not None else range(0, n)
    nq = len(r)
    dev = point_cloud.device
    normals = torch.zeros((max(1, nq), 3), dtype=torch.float64, device=dev)
    curv = torch.zeros(max(1, nq), dtype=torch.float64, device=dev)
    d = point_cloud.desc()
    check(lib().pb200_compute_normals_range(ctx._h, C.byref(d), k_nn, r.start, nq, C.c_void_p(normals.data_ptr()),
                                            C.c_void_p(curv.data_ptr())))
    return normals[:nq], curv[:nq]


# ---- segmentation (pasture-algorithms/src/segmentation.rs) --------------------------------------------------
RANSAC_PLANE, RANSAC_LINE = 0, 1


class Plane:
    """segmentation.rs:19-28: ax + by + cz + d = 0 and the number of inliers"""

    def __init__(self, a, b, c, d, ranking):
        self.a, self.b, self.c, self.d, self.ranking = a, b, c, d, ranking

    def __repr__(self):
        return f"Plane {{ a: {self.a}, b: {self.b}, c: {self.c}, d: {self.d}, ranking: {self.ranking} }}"


class Line:
    """segmentation.rs:10-17"""

    def __init__(self, first, second, ranking):
        self.first, self.second, self.ranking = first, second, ranking

    def __repr__(self):
        return f"Line {{ first: {self.first}, second: {self.second}, ranking: {self.ranking} }}"


def ransac_rank_samples(buffer, kind, samples, distance_threshold, ctx=None):
    """generate_{plane,line}_model (:98-138) for given draws -> (models [m, 4|6] f64, rankings [m] u64) numpy"""
    ctx = context_for(ctx, buffer)
    s = np.ascontiguousarray(samples, dtype=np.uint64)
    m = s.shape[0]
    models = np.zeros((m, 4 if kind == RANSAC_PLANE else 6), dtype=np.float64)
    ranks = np.zeros(m, dtype=np.uint64)
    d = buffer.desc()
    check(lib().pb200_ransac_rank_samples(ctx._h, C.byref(d), kind, s.ctypes.data, m, float(distance_threshold),
                                          models.ctypes.data, ranks.ctypes.data))
    return models, ranks


def ransac_rank_models(buffer, kind, models, distance_threshold, ctx=None):
    ctx = context_for(ctx, buffer)
    mm = np.ascontiguousarray(models, dtype=np.float64)
    ranks = np.zeros(mm.shape[0], dtype=np.uint64)
    d = buffer.desc()
    check(lib().pb200_ransac_rank_models(ctx._h, C.byref(d), kind, mm.ctypes.data, mm.shape[0], float(distance_threshold),
                                         ranks.ctypes.data))
    return ranks


def ransac_inliers(buffer, kind, model, distance_threshold, ctx=None):
    """indices (int64 tensor on the buffer's device, ascending) of the points within the threshold of `model`"""
    ctx = context_for(ctx, buffer)
    mm = np.ascontiguousarray(model, dtype=np.float64)
    n = buffer.len()
    idx = torch.zeros(max(1, n), dtype=torch.int64, device=buffer.device)
    cnt = C.c_uint64(0)
    d = buffer.desc()
    check(lib().pb200_ransac_inliers(ctx._h, C.byref(d), kind, mm.ctypes.data, float(distance_threshold),
                                     C.c_void_p(idx.data_ptr()), n, C.byref(cnt)))
    return idx[: int(cnt.value)]


def _ransac(buffer, kind, distance_threshold, num_of_iterations, seed, ctx):
    ctx = context_for(ctx, buffer)
    if seed is None:  # the reference draws from thread_rng(): every call tries different models
        seed = int.from_bytes(__import__("os").urandom(8), "little")
    n = buffer.len()
    model = np.zeros(4 if kind == RANSAC_PLANE else 6, dtype=np.float64)
    rank = C.c_uint64(0)
    idx = torch.zeros(max(1, n), dtype=torch.int64, device=buffer.device)
    d = buffer.desc()
    check(lib().pb200_ransac(ctx._h, C.byref(d), kind, float(distance_threshold), int(num_of_iterations), int(seed),
                             model.ctypes.data, C.byref(rank), C.c_void_p(idx.data_ptr()), n))
    r = int(rank.value)
    if kind == RANSAC_PLANE:
        return Plane(*model.tolist(), r), idx[:r]
    return Line(model[:3].copy(), model[3:].copy(), r), idx[:r]


def ransac_plane_par(buffer, distance_threshold, num_of_iterations, seed=None, ctx=None):  # segmentation.rs:180
    return _ransac(buffer, RANSAC_PLANE, distance_threshold, num_of_iterations, seed, ctx)


def ransac_line_par(buffer, distance_threshold, num_of_iterations, seed=None, ctx=None):  # :291
    return _ransac(buffer, RANSAC_LINE, distance_threshold, num_of_iterations, seed, ctx)


ransac_plane_serial = ransac_plane_par  # :240 -- one implementation: all iterations are ranked in the same pass
ransac_line_serial = ransac_line_par  # :350


class Projection:
    """reprojection.rs:10-70 with an enumerated operation pipeline instead of a PROJ string"""

    def __init__(self, source_crs, target_crs):
        ops = (ProjOp * 8)()
        n = check(lib().pb200_proj_pipeline_for_crs(source_crs.encode(), target_crs.encode(), ops, 8))
        self.ops, self.n_ops = ops, n


def reproject_point_cloud_within(point_cloud, source_crs, target_crs, ctx=None):  # reprojection.rs:132-146
    ctx = context_for(ctx, point_cloud)
    p = Projection(source_crs, target_crs)
    d = point_cloud.desc()
    check(lib().pb200_reproject(ctx._h, C.byref(d), None, p.ops, p.n_ops))


def reproject_point_cloud_between(source_point_cloud, target_point_cloud, source_crs, target_crs, ctx=None):  # :201-227
    ctx = context_for(ctx, source_point_cloud)
    p = Projection(source_crs, target_crs)
    sd, dd = source_point_cloud.desc(), target_point_cloud.desc()
    check(lib().pb200_reproject(ctx._h, C.byref(sd), C.byref(dd), p.ops, p.n_ops))


def synth_las_fmt0_records(n, first_index=0, seed=42, device="cuda", ctx=None):
    """C2/C5 input stream generated in HBM: VectorBuffer of raw LAS format-0 records"""
    from .layout import PointLayout
    ctx = ctx.bind_current_stream() if ctx is not None else get_context(device)
    buf = VectorBuffer(PointLayout.las_raw(0), n, device)
    check(lib().pb200_synth_las_fmt0_records(ctx._h, C.c_void_p(buf.data.data_ptr()), first_index, n, seed))
    return buf


def synth_terrain_positions(n, first_index=0, seed=42, device="cuda", ctx=None):
    """C3/C4 input stream: HashMapBuffer with one packed Vec3f64 POSITION_3D column"""
    from .layout import PointLayout, attributes
    ctx = ctx.bind_current_stream() if ctx is not None else get_context(device)
    buf = HashMapBuffer(PointLayout.from_attributes([attributes.POSITION_3D]), n, device)
    check(lib().pb200_synth_terrain_positions(ctx._h, C.c_void_p(buf.columns[0].data_ptr()), first_index, n, seed))
    return buf
