"""Point-range sharding across GPUs: one process per GPU, rank r owns the contiguous range
[r*N/G, (r+1)*N/G) -- the `convert_into_range` / `slice(range)` contract of the reference
(buffer_conversion.rs:292, containers/slice.rs:16-43).  The only exchange on the conversion path is the global
AABB: one all-reduce(MIN) of [min xyz, -max xyz] (48 bytes).  The voxel-grid filter adds one key-range all-to-all of
per-shard partial sums (`voxelgrid_filter_sharded`)."""
import ctypes as C

import torch
import torch.distributed as dist

F64_MAX = 1.7976931348623157e308


def shard_range(n_points, rank, world_size):
    """contiguous, balanced: the first n % world ranks hold one extra point"""
    base, extra = divmod(int(n_points), int(world_size))
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return range(begin, end)


def pack_bounds(aabb_or_none, device="cpu"):
    """AABB (or None for an empty shard) -> the 6-vector the convert kernel produces: [min xyz, -max xyz]"""
    if aabb_or_none is None:
        return torch.full((6,), F64_MAX, dtype=torch.float64, device=device)
    mn, mx = aabb_or_none
    return torch.tensor(list(mn) + [-v for v in mx], dtype=torch.float64, device=device)


def unpack_bounds(minmax6):
    """[min xyz, -max xyz] -> (min, max) or None if no shard contributed a point"""
    v = minmax6.detach().cpu().tolist()
    if v[0] == F64_MAX and v[3] == F64_MAX:
        return None
    return tuple(v[:3]), tuple(-x for x in v[3:])


def allreduce_bounds(minmax6, group=None):
    """AABB::union over all shards (math/bounds.rs:109-122) as ONE collective: min over [min, -max]"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(minmax6, op=dist.ReduceOp.MIN, group=group)
    return minmax6


# ---- sharded voxel grid: partials -> key-range all-to-all -> merge --------------------------------------

def key_range_boundaries(cells_x, bits_y, bits_z, world_size):
    """Rank d finalises the voxels whose x index lies in [d*cells_x//W, (d+1)*cells_x//W): returns the W-1 packed keys at
    which ownership changes (a key k belongs to rank = number of boundaries <= k)."""
    cells_x = max(1, int(cells_x))
    return [((d * cells_x) // world_size) << (bits_y + bits_z) for d in range(1, world_size)]


def split_sizes(sorted_keys, boundaries):
    """how many entries of an ascending key tensor go to each destination rank"""
    if not boundaries:
        return [int(sorted_keys.numel())]
    b = torch.tensor(boundaries, dtype=sorted_keys.dtype, device=sorted_keys.device)
    cut = torch.searchsorted(sorted_keys, b, right=False).tolist()
    edges = [0] + cut + [int(sorted_keys.numel())]
    return [edges[i + 1] - edges[i] for i in range(len(edges) - 1)]


def exchange_partials(keys, counts, sums, send_sizes, group=None):
    """all-to-all of the three partial arrays; what arrives is concatenated in source-rank order.  Works on CUDA tensors
    (NCCL) and on CPU tensors (gloo); without a process group it is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys, counts, sums
    world = dist.get_world_size(group)
    assert len(send_sizes) == world
    dev = keys.device
    sent = torch.tensor(send_sizes, dtype=torch.int64, device=dev)
    recv = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, sent, group=group)
    recv_sizes = recv.tolist()
    m = sum(recv_sizes)
    out_keys = torch.empty(m, dtype=keys.dtype, device=dev)
    out_counts = torch.empty(m, dtype=counts.dtype, device=dev)
    out_sums = torch.empty((m, 3), dtype=sums.dtype, device=dev)
    dist.all_to_all_single(out_keys, keys.contiguous(), recv_sizes, list(send_sizes), group=group)
    dist.all_to_all_single(out_counts, counts.contiguous(), recv_sizes, list(send_sizes), group=group)
    dist.all_to_all_single(out_sums, sums.contiguous(), recv_sizes, list(send_sizes), group=group)
    return out_keys, out_counts, out_sums


def voxelgrid_filter_sharded(shard, leafsize_x, leafsize_y, leafsize_z, group=None, ctx=None):
    """voxel_grid.rs:109-165 over a cloud sharded by point range, one shard per rank.  Returns (VoxelPartials, centroids)
    for THIS rank's key range (ranks own ascending, disjoint x-index ranges, so concatenating the ranks' results in rank
    order gives the reference's output order), or None if the whole cloud is empty.
    Collectives: all-reduce(MIN) of the bounds (48 B), all-to-all of the send counts, 3 x all-to-all of the partials."""
    from . import algorithms as alg
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    b = alg.calculate_bounds(shard) if shard.len() else None
    v = pack_bounds(None if b is None else (b.min(), b.max()), shard.device)
    g = unpack_bounds(allreduce_bounds(v, group))
    if g is None:
        return None
    part = alg.voxelgrid_partials(shard, leafsize_x, leafsize_y, leafsize_z, g, ctx)
    bounds = key_range_boundaries(part.cells[0], part.bits[1], part.bits[2], world)
    keys, counts, sums = exchange_partials(part.keys, part.counts, part.sums, split_sizes(part.keys, bounds), group)
    return alg.voxelgrid_merge_partials(keys, counts, sums, part.bits, part.cells, ctx)


def balanced_key_boundaries(sorted_keys, world_size, group=None, samples=2048):
    """W - 1 splitter keys such that every rank finalises about the same number of voxels WHATEVER the cloud looks like:
    each rank contributes `samples` evenly spaced quantiles of its own (ascending) partial keys, the gathered samples are
    sorted and cut into W equal parts.  (Equal x-index ranges, the round-1 rule, leave most ranks idle for any cloud that
    does not fill its bounding box evenly.)  A key k belongs to rank = number of splitters <= k."""
    have = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have else 1
    if world_size <= 1:
        return []
    m = int(sorted_keys.numel())
    dev = sorted_keys.device
    if m:
        pos = (torch.arange(samples, device=dev, dtype=torch.int64) * (m - 1)) // max(1, samples - 1)  # exact integer quantiles
        mine = sorted_keys[pos]
    else:
        mine = torch.full((samples,), -1, dtype=sorted_keys.dtype, device=dev)  # -1: no contribution
    if world > 1:
        allk = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine, group=group)
        allk = torch.cat(allk)
    else:
        allk = mine
    allk = allk[allk >= 0]
    if allk.numel() == 0:
        return [0] * (world_size - 1)
    allk = torch.sort(allk).values
    cuts = [int(allk[min(allk.numel() - 1, (d * allk.numel()) // world_size)].item()) for d in range(1, world_size)]
    return cuts


def _all_to_all_rows(t, send_sizes, recv_sizes, group):
    out = torch.empty((sum(recv_sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out, t.contiguous(), list(recv_sizes), list(send_sizes), group=group)
    return out


def _exchange_sizes(send_sizes, device, group):
    sent = torch.tensor(send_sizes, dtype=torch.int64, device=device)
    recv = torch.empty_like(sent)
    dist.all_to_all_single(recv, sent, group=group)
    return recv.tolist()


def voxelgrid_filter_sharded_layout(shard, leafsize_x, leafsize_y, leafsize_z, filtered_layout=None, group=None, ctx=None,
                                    return_keys=False, timings=None):
    """voxel_grid.rs:109-165 with ALL attribute reductions over a cloud sharded by point range (one shard per rank):
    global AABB (all-reduce) -> per-shard partials for every attribute of `filtered_layout` -> balanced key-range
    boundaries from sampled keys -> ONE set of key-range all-to-alls (voxel rows, then one run list per mode attribute)
    -> merge.  Returns this rank's part of the filtered cloud (a HashMapBuffer; concatenating the ranks' parts in rank
    order gives the reference's output order) or None for an empty cloud.  `timings` (a dict) receives per-phase
    milliseconds (device-synchronised) when given."""
    import time
    from . import algorithms as alg
    have = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have else 1
    layout = filtered_layout or shard.point_layout()

    def mark(name, t0):
        if timings is not None:
            torch.cuda.synchronize(shard.device)
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t0 = time.perf_counter()
    b = alg.calculate_bounds(shard) if shard.len() else None
    v = pack_bounds(None if b is None else (b.min(), b.max()), shard.device)
    g = unpack_bounds(allreduce_bounds(v, group))
    t0 = mark("bounds+allreduce", t0)
    if g is None:
        return None
    part = alg.voxelgrid_partials_layout(shard, leafsize_x, leafsize_y, leafsize_z, g, layout, ctx)
    t0 = mark("partials", t0)
    if world > 1:
        bounds = balanced_key_boundaries(part.pos.keys, world, group)
        send = split_sizes(part.pos.keys, bounds)
        recv = _exchange_sizes(send, shard.device, group)
        pos = part.pos
        keys = _all_to_all_rows(pos.keys, send, recv, group)
        counts = _all_to_all_rows(pos.counts, send, recv, group)
        sums = _all_to_all_rows(pos.sums, send, recv, group)
        cols = _all_to_all_rows(part.columns, send, recv, group) if part.columns.shape[1] else torch.empty((keys.numel(), 0), dtype=torch.float64, device=shard.device)
        modes = []
        for mk, mc in part.modes:  # run lists are ascending in (voxel key << 16 | value): the same splitters, shifted
            msend = split_sizes(mk, [bk << 16 for bk in bounds])
            mrecv = _exchange_sizes(msend, shard.device, group)
            modes.append((_all_to_all_rows(mk, msend, mrecv, group), _all_to_all_rows(mc, msend, mrecv, group)))
        part = alg.VoxelAttrPartials(alg.VoxelPartials(keys, counts, sums, pos.bits, pos.cells), cols, part.column_is_max, modes)
        t0 = mark("all_to_all", t0)
    out = alg.voxelgrid_merge_partials_layout(part, layout, ctx, return_keys=return_keys)
    mark("merge", t0)
    return out


# ---- minmax_attribute over point-range shards ("minmax ints: same with the attribute's dtype") --------------

def combine_minmax(local, first_value, attribute, group=None, device="cpu"):
    """The exchange step of minmax_attribute_sharded (host logic, any backend).  `local` = this rank's NaN-ignoring
    (min, max) in the attribute's dtype or None for an empty shard, `first_value` = the attribute of the shard's first point
    (None for an empty shard).  ONE all-reduce(MIN) over [min, -max] (floats, as f64) or over the order-preserving int64
    images of the values (integers of every width, u64 included) plus one small all-gather for the seed rule."""
    import numpy as np
    from .containers import _NP, _VEC3
    have = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have else 1
    dt = attribute.datatype()
    comp = _NP.get(dt) or _NP[_VEC3[dt]]
    nc = 1 if dt in _NP else 3
    is_float = np.issubdtype(comp, np.floating)
    flip = comp == np.uint64  # u64 -> int64 with the sign bit flipped: order preserving

    def enc(v):
        v = np.atleast_1d(np.asarray(v, dtype=comp))
        return (v ^ np.uint64(1 << 63)).view(np.int64) if flip else v.astype(np.int64)

    def dec(v):
        return (v.view(np.uint64) ^ np.uint64(1 << 63)).astype(comp) if flip else v.astype(comp)
    if is_float:
        vec = np.full(2 * nc, F64_MAX)
        if local is not None:
            mn, mx = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in local)
            vec = np.concatenate([mn, -mx])
        t = torch.tensor(vec, dtype=torch.float64, device=device)
    else:
        big = np.iinfo(np.int64).max
        vec = np.full(2 * nc, big, np.int64)
        if local is not None:  # the max travels as ~v = -(v + 1): no overflow at the extremes, one MIN reduces both
            vec = np.concatenate([enc(local[0]), ~enc(local[1])])
        t = torch.tensor(vec, dtype=torch.int64, device=device)
    seed_nan = [0.0] * nc
    if first_value is not None and is_float:
        seed_nan = [1.0 if np.isnan(x) else 0.0 for x in np.atleast_1d(np.asarray(first_value, dtype=np.float64)).reshape(nc)]
    meta = torch.tensor([0.0 if first_value is None else 1.0] + seed_nan, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
    else:
        metas = [meta]
    metas = [m.cpu().numpy() for m in metas]
    first = next((m for m in metas if m[0] > 0), None)  # rank order = point order: the first non-empty shard holds point 0
    if first is None:
        return None  # minmax.rs:14-16
    r = t.cpu().numpy()
    if is_float:
        mn, mx = r[:nc].copy(), -r[nc:]
        for c in range(nc):
            if first[1 + c] > 0:  # a NaN seed is never replaced (math/minmax.rs:62-96)
                mn[c] = mx[c] = np.nan
        mn, mx = mn.astype(comp), mx.astype(comp)
    else:
        mn, mx = dec(r[:nc].copy()), dec(~r[nc:])
    return (mn[0], mx[0]) if nc == 1 else (mn, mx)


def minmax_attribute_sharded(buffer, attribute, group=None, ctx=None):
    """minmax_attribute::<T> (pasture-algorithms/src/minmax.rs:13-51) of a cloud sharded by point range: every rank folds its
    shard on its GPU (NaN ignored: pb200_minmax_attribute_partial) and `combine_minmax` exchanges the partial results and
    restores the reference's seed rule -- (min, max) start from the cloud's first value and a NaN seed is never replaced --
    from the first point of the first non-empty shard.  Returns (min, max) in the attribute's dtype on every rank, or None
    for an empty cloud."""
    from . import algorithms as alg
    have = dist.is_available() and dist.is_initialized()
    on_gpu = have and dist.get_backend(group) == "nccl"
    local = alg.minmax_attribute(buffer, attribute, ctx=ctx, partial=True) if buffer.len() else None
    first = buffer.slice_first(attribute) if buffer.len() else None
    return combine_minmax(local, first, attribute, group, buffer.device if on_gpu else "cpu")


# ---- kNN / normals: replicas only -------------------------------------------------------------------------

def compute_normals_sharded(point_cloud, k_nn, group=None, gather=False, ctx=None):
    """compute_normals (normal_estimation.rs:79-130) over the ranks of `group`: every rank holds the WHOLE position column
    (a replica: 2.4 GB per 100 M points), builds the same tree and answers the queries of its own point range
    (`shard_range`), the `for point in points` loop of :106-127 cut into contiguous pieces.  No exchange is needed for the
    computation; gather=True all-gathers the slices so that every rank ends up with all normals / curvatures.
    Returns (normals, curvature, range) -- the rank's slice and the point range it covers (or the full arrays)."""
    from . import algorithms as alg
    have = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have else 1
    rank = dist.get_rank(group) if have else 0
    n = point_cloud.len()
    r = shard_range(n, rank, world)
    normals, curv = alg.compute_normals(point_cloud, k_nn, query_range=r, ctx=ctx)
    if not gather or world == 1:
        return normals, curv, r
    sizes = [len(shard_range(n, q, world)) for q in range(world)]
    dev = normals.device
    on_gpu = dist.get_backend(group) == "nccl"
    outs_n = [torch.empty((sz, 3), dtype=torch.float64, device=dev if on_gpu else "cpu") for sz in sizes]
    outs_c = [torch.empty(sz, dtype=torch.float64, device=dev if on_gpu else "cpu") for sz in sizes]
    dist.all_gather(outs_n, normals if on_gpu else normals.cpu(), group=group)
    dist.all_gather(outs_c, curv if on_gpu else curv.cpu(), group=group)
    return torch.cat(outs_n).to(dev), torch.cat(outs_c).to(dev), range(0, n)


# ---- peer-memory communicator: the global AABB inside the convert kernel (no NCCL on the data path) ----------------

class PeerComm:
    """pb200_comm: every rank owns a small exchange buffer that its peers map over NVLink.
    `PeerComm(ctx, group=None)` -- one process per GPU: the CUDA-IPC handles travel through ONE host-side all-gather at
    set-up time (torch.distributed, any backend); after that `convert_into_range_with_global_bounds` needs no
    collective library call.  `PeerComm.local_group(contexts)` wires several contexts of one process by raw pointers."""

    HANDLE_BYTES = 64

    def __init__(self, ctx, group=None, _rank=None, _world=None, _connect=True):
        from ._lib import check, lib
        self.ctx = ctx
        self._h = None
        have = dist.is_available() and dist.is_initialized()
        self.rank = _rank if _rank is not None else (dist.get_rank(group) if have else 0)
        self.world = _world if _world is not None else (dist.get_world_size(group) if have else 1)
        collective = _connect and self.world > 1
        # Every rank takes part in every collective of the set-up, whatever happens locally: a rank that fails must not
        # leave its peers waiting.  Failures are agreed on at the end (all-reduce MIN of an ok flag) and raised everywhere.
        error = None
        mine = (C.c_uint8 * self.HANDLE_BYTES)()
        try:
            h = C.c_void_p()
            check(lib().pb200_comm_create(ctx._h, self.rank, self.world, C.byref(h)))
            self._h = h
            if collective:
                check(lib().pb200_comm_handle(self._h, mine))
        except Exception as exc:  # noqa: BLE001
            error = exc
        if not collective:
            if error is not None:
                raise error
            return
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(mine) if error is None else b"", group=group)
        if error is None and all(len(g) == self.HANDLE_BYTES for g in gathered):
            try:
                blob = (C.c_uint8 * (self.HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(gathered))
                check(lib().pb200_comm_connect(self._h, blob))
            except Exception as exc:  # noqa: BLE001
                error = exc
        elif error is None:
            error = RuntimeError("a peer could not create its exchange buffer")
        on_gpu = dist.get_backend(group) == "nccl"
        flag = torch.tensor([0 if error is not None else 1], dtype=torch.int32,
                            device=torch.device("cuda", ctx.device) if on_gpu else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # also the barrier: nobody publishes before all mappings exist
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError(f"peer-memory communicator set-up failed on at least one rank (this rank: {error})")

    @classmethod
    def local_group(cls, contexts):
        """one process, several contexts (streams / devices): ranks are the list positions"""
        import os
        from ._lib import check, lib
        per_device = {}
        for c in contexts:
            per_device[c.device] = per_device.get(c.device, 0) + 1
        crowd = max(per_device.values()) if per_device else 0
        if crowd > 1:
            # the fused kernels of these ranks spin-wait for each other INSIDE the kernel: they must run concurrently, so
            # their streams must not share a hardware work queue (default: 8 connections, streams alias beyond that)
            have = int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8") or 8)
            if have < 2 * crowd:
                raise RuntimeError(f"{crowd} logical ranks share one device: set CUDA_DEVICE_MAX_CONNECTIONS >= {2 * crowd} "
                                   f"(now {have}) before CUDA is initialised, or the ranks' kernels can serialise and time out")
        comms = [cls(c, _rank=r, _world=len(contexts), _connect=False) for r, c in enumerate(contexts)]
        ptrs = (C.c_void_p * len(comms))()
        for r, cm in enumerate(comms):
            p = C.c_void_p()
            check(lib().pb200_comm_exchange_ptr(cm._h, C.byref(p)))
            ptrs[r] = p.value
        for cm in comms:
            check(lib().pb200_comm_connect_ptrs(cm._h, ptrs))
        return comms

    def check(self):
        """synchronise and raise if a peer never arrived (the kernel gives up after 10 s)"""
        from ._lib import check, lib
        check(lib().pb200_comm_check(self._h))

    def close(self):
        from ._lib import lib
        if self._h:
            lib().pb200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
