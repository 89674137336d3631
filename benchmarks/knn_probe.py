"""kNN / normals probe: per-phase timings and traversal statistics for the C4 stream (diagnostics, not a bench line)"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb
from pasture_b200 import algorithms as alg
from pasture_b200.context import get_context

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=20_000_000)
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--radii", default="-1")
ap.add_argument("--per-axis", type=int, default=0)
ap.add_argument("--stats", type=int, default=1)
ap.add_argument("--heap", type=int, default=0)
args = ap.parse_args()
ctx = get_context()
src = alg.synth_terrain_positions(args.points)
ctx.set_param("knn.per_axis_codes", args.per_axis)
ctx.set_param("knn.heap", args.heap)
for r in [int(x) for x in args.radii.split(",")]:
    ctx.set_param("knn.init_radius", r)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        normals, curv = alg.compute_normals(src, args.k)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print(f"init_radius={r} per_axis={args.per_axis} normals k={args.k}: {ms:.1f} ms", flush=True)
    if args.stats:
        ctx.set_param("knn.stats", 1)
        idx = alg.knn(src, args.k, with_distances=False)
        torch.cuda.synchronize()
        st = idx.reshape(-1)[: 3 * args.points].view(args.points, 3).to(torch.float64)
        print("   mean nodes/buckets/offers per query:", [round(x, 2) for x in st.mean(0).tolist()], "max:", st.max(0).values.tolist(), flush=True)
        ctx.set_param("knn.stats", 0)
