#!/usr/bin/env python
"""Sharded voxel-grid filter over N GPUs (SURVEY 8e, C3 stream): one process per GPU under torchrun, rank r owns point
indices [r*P, (r+1)*P) of the synthetic terrain.  Per step: local AABB -> all-reduce(MIN) -> per-shard partials on the
global grid -> key-range all-to-all (NCCL) -> merge.  Prints one JSON line (informational; the headline is bench.py).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \\
      benchmarks/sharded_voxel.py --points-per-gpu 50000000 --check
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb  # noqa: E402
from pasture_b200 import algorithms as alg, sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points-per-gpu", type=int, default=50_000_000)
    ap.add_argument("--leaf", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare with the single-device filter on rank 0 (needs N*P points on one GPU)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = args.points_per_gpu
    shard = alg.synth_terrain_positions(p, first_index=rank * p)
    best = None
    for step in range(args.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        part, cent = sharding.voxelgrid_filter_sharded(shard, args.leaf, args.leaf, args.leaf)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if step > 0:
            best = float(ms) if best is None else min(best, float(ms))
    tot = torch.tensor([part.len(), int(part.counts.sum())], dtype=torch.int64, device="cuda")
    first_last = torch.tensor([int(part.keys[0]) if part.len() else -1, int(part.keys[-1]) if part.len() else -1], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
        fl = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(fl, first_last)
    else:
        fl = [first_last]
    ordered = all(int(fl[i][1]) < int(fl[i + 1][0]) for i in range(world - 1) if int(fl[i][1]) >= 0 and int(fl[i + 1][0]) >= 0)
    assert bool((part.keys[1:] > part.keys[:-1]).all()) and ordered, "voxel keys are not globally ascending"
    assert int(tot[1]) == p * world, "points lost in the merge"
    check = None
    if args.check:
        # gather everything on rank 0 and compare with the single-device filter over the whole cloud
        sizes = torch.tensor([part.len()], dtype=torch.int64, device="cuda")
        all_sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        if world > 1:
            dist.all_gather(all_sizes, sizes)
        else:
            all_sizes = [sizes]
        mx = int(max(int(s) for s in all_sizes))
        pad_k = torch.zeros(mx, dtype=torch.int64, device="cuda"); pad_k[: part.len()] = part.keys
        pad_c = torch.zeros((mx, 3), dtype=torch.float64, device="cuda"); pad_c[: part.len()] = cent
        gk = [torch.zeros_like(pad_k) for _ in range(world)]
        gc = [torch.zeros_like(pad_c) for _ in range(world)]
        if world > 1:
            dist.all_gather(gk, pad_k); dist.all_gather(gc, pad_c)
        else:
            gk, gc = [pad_k], [pad_c]
        if rank == 0:
            keys = torch.cat([gk[r][: int(all_sizes[r])] for r in range(world)])
            cents = torch.cat([gc[r][: int(all_sizes[r])] for r in range(world)])
            whole = alg.synth_terrain_positions(p * world)
            single, skeys = alg.voxelgrid_filter(whole, args.leaf, args.leaf, args.leaf, return_keys=True)
            by, bz = part.bits[1], part.bits[2]
            import numpy as np
            packed = (skeys[:, 0].astype(np.int64) << (by + bz)) | (skeys[:, 1].astype(np.int64) << bz) | skeys[:, 2].astype(np.int64)
            same_keys = bool(np.array_equal(keys.cpu().numpy(), packed))
            ref = torch.from_numpy(single.view_attribute("Position3D")).cuda() if not torch.is_tensor(single.view_attribute("Position3D")) else single.view_attribute("Position3D")
            rel = float(((cents - ref).abs() / ref.abs().clamp_min(1e-300)).max()) if same_keys else float("nan")
            check = {"voxel_keys_identical": same_keys, "centroid_max_rel_err": rel, "tolerance": 1e-9}
            assert same_keys and rel <= 1e-9, check
    if rank == 0:
        print(json.dumps({"config": "sharded voxel-grid filter (C3 stream): bounds all-reduce + partials + key-range all-to-all + merge",
                          "n_gpus": world, "points_per_gpu": p, "leaf": args.leaf, "voxels": int(tot[0]), "ms_per_step": best,
                          "points_per_s": p * world / (best * 1e-3), "check": check}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
