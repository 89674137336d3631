#!/usr/bin/env python
"""Sharded voxel-grid filter over N GPUs (SURVEY 8e, C3 stream): one process per GPU under torchrun, rank r owns point
indices [r*P, (r+1)*P) of the synthetic terrain (+ synthetic LAS attributes with --attributes las).  Per step:
local AABB -> all-reduce(MIN) -> per-shard partials of every attribute on the global grid -> balanced key-range
boundaries from sampled keys -> key-range all-to-alls (NCCL) -> merge.  Prints one JSON line with the step time
(max over ranks), a per-phase table (one extra step with a device synchronisation after every phase) and, with --check,
the comparison against the single-device filter over the WHOLE cloud on rank 0 (informational; the headline is bench.py).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \\
      benchmarks/sharded_voxel.py --points-per-gpu 50000000 --check
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb  # noqa: E402
from pasture_b200 import algorithms as alg, sharding  # noqa: E402
from pasture_b200 import attributes as A  # noqa: E402


def make_shard(p, rank, with_attrs):
    pos = alg.synth_terrain_positions(p, first_index=rank * p)
    if not with_attrs:
        return pos, pos.point_layout()
    layout = pb.PointLayout.from_attributes([A.POSITION_3D, A.INTENSITY, A.CLASSIFICATION, A.RETURN_NUMBER, A.GPS_TIME])
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    cols = [pos.columns[0],
            torch.randint(0, 256, (2 * p,), dtype=torch.uint8, device="cuda", generator=g),           # Intensity: mean (u16)
            torch.randint(0, 6, (p,), dtype=torch.uint8, device="cuda", generator=g),                 # Classification: mode
            torch.randint(0, 8, (p,), dtype=torch.uint8, device="cuda", generator=g),                 # ReturnNumber: mode
            (torch.rand(p, dtype=torch.float64, device="cuda", generator=g) * 1e5).view(torch.uint8)]  # GpsTime: max-pool
    return pb.HashMapBuffer(layout, p, "cuda", columns=cols), layout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points-per-gpu", type=int, default=50_000_000)
    ap.add_argument("--leaf", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--attributes", default="position", choices=["position", "las"])
    ap.add_argument("--check", action="store_true", help="compare with the single-device filter on rank 0 (needs N*P points on one GPU)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = args.points_per_gpu
    shard, layout = make_shard(p, rank, args.attributes == "las")
    leaf = (args.leaf,) * 3
    best, out, keys = None, None, None
    for step in range(args.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sharding.voxelgrid_filter_sharded_layout(shard, *leaf, filtered_layout=layout)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if step > 0:
            best = float(ms) if best is None else min(best, float(ms))
    phases = {}
    if world > 1:
        dist.barrier()
    out = sharding.voxelgrid_filter_sharded_layout(shard, *leaf, filtered_layout=layout, timings=phases)
    if args.check:  # (the (ix, iy, iz) keys are unpacked on the host: kept out of the timed phases)
        out, keys = sharding.voxelgrid_filter_sharded_layout(shard, *leaf, filtered_layout=layout, return_keys=True)
    ph = torch.tensor([phases.get(k, 0.0) for k in ("bounds+allreduce", "partials", "all_to_all", "merge")], dtype=torch.float64, device="cuda")
    ph_max = ph.clone()
    if world > 1:
        dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
    v_local = torch.tensor([out.len()], dtype=torch.int64, device="cuda")
    v_all = [torch.zeros_like(v_local) for _ in range(world)]
    if world > 1:
        dist.all_gather(v_all, v_local)
    else:
        v_all = [v_local]
    sizes = [int(x) for x in v_all]
    check = None
    if args.check:
        # every rank's result and shard go to rank 0, which runs the single-device filter over the whole cloud
        mx = max(sizes)
        names = [layout.at(i).name() for i in range(len(layout))]
        gathered = []
        for i, nm in enumerate(names):
            sz = layout.at(i).size()
            pad = torch.zeros(mx * sz, dtype=torch.uint8, device="cuda")
            pad[: out.len() * sz] = out.columns[i][: out.len() * sz]
            g = [torch.zeros_like(pad) for _ in range(world)]
            if world > 1:
                dist.all_gather(g, pad)
            else:
                g = [pad]
            gathered.append(torch.cat([g[r][: sizes[r] * sz] for r in range(world)]))
        kp = torch.zeros((mx, 3), dtype=torch.int64, device="cuda")
        kp[: out.len()] = torch.from_numpy(keys.astype(np.int64)).cuda()
        gk = [torch.zeros_like(kp) for _ in range(world)]
        if world > 1:
            dist.all_gather(gk, kp)
        else:
            gk = [kp]
        all_keys = torch.cat([gk[r][: sizes[r]] for r in range(world)])
        src_cols = []
        for i in range(len(layout)):
            sz = layout.at(i).size()
            g = [torch.zeros(p * sz, dtype=torch.uint8, device="cuda") for _ in range(world)]
            if world > 1:
                dist.all_gather(g, shard.columns[i][: p * sz].contiguous())
            else:
                g = [shard.columns[i][: p * sz]]
            src_cols.append(torch.cat(g))
        if rank == 0:
            whole = pb.HashMapBuffer(layout, p * world, "cuda", columns=src_cols)
            single, skeys = alg.voxelgrid_filter(whole, *leaf, filtered_layout=layout, return_keys=True)
            same_keys = single.len() == sum(sizes) and bool(np.array_equal(all_keys.cpu().numpy(), skeys.astype(np.int64)))
            check = {"voxel_keys_identical": same_keys, "attributes": {}}
            ok = same_keys
            for i, nm in enumerate(names):
                sz = layout.at(i).size()
                a, b = gathered[i][: single.len() * sz], single.columns[i][: single.len() * sz]
                if layout.at(i).datatype() == pb.PointAttributeDataType.Vec3f64 and same_keys:
                    av, bv = a.view(torch.float64), b.view(torch.float64)
                    rel = float(((av - bv).abs() / bv.abs().clamp_min(1e-300)).max())
                    check["attributes"][nm] = {"max_rel_err": rel, "tolerance": 1e-9}
                    ok = ok and rel <= 1e-9
                elif same_keys:
                    eq = bool(torch.equal(a, b))
                    check["attributes"][nm] = {"identical": eq}
                    ok = ok and eq
            assert ok, check
    if rank == 0:
        tot = sum(sizes)
        print(json.dumps({"config": "sharded voxel-grid filter (C3 stream): bounds all-reduce + per-attribute partials + balanced key-range "
                                    "all-to-all + merge", "attributes": [layout.at(i).name() for i in range(len(layout))],
                          "n_gpus": world, "points_per_gpu": p, "leaf": args.leaf, "voxels": tot, "voxels_per_rank": sizes,
                          "ms_per_step": best, "points_per_s": p * world / (best * 1e-3),
                          "phase_ms_max_over_ranks": dict(zip(("bounds+allreduce", "partials", "all_to_all", "merge"), [round(float(x), 3) for x in ph_max])),
                          "phase_note": "one extra step with a device synchronisation after every phase (their sum exceeds ms_per_step)",
                          "check": check}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
