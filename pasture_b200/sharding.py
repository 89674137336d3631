"""Point-range sharding across GPUs: one process per GPU, rank r owns the contiguous range
[r*N/G, (r+1)*N/G) -- the `convert_into_range` / `slice(range)` contract of the reference
(buffer_conversion.rs:292, containers/slice.rs:16-43).  The only exchange on the conversion path is the global
AABB: one all-reduce(MIN) of [min xyz, -max xyz] (48 bytes)."""
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
