#!/usr/bin/env python
"""prof_all.py -- launches every kernel family of libpasture_b200 once on a representative workload so that ONE
`ncu --set full` run captures them all (the per-kernel evidence the north star asks for):

  ncu --set full --clock-control none --import-source on -k regex:'^(convert|bounds|minmax|morton|voxel|radix|heads|lbvh|gather_pos|reproject|filter_c|ransac_r|return_hist|pnts)' \\
      -c 80 -o gpurun_out/prof_all_r2 python benchmarks/prof_all.py --points 50000000
  python benchmarks/ncu_summary.py gpurun_out/prof_all_r2.ncu-rep --all --bytes-json gpurun_out/prof_all_bytes.json --out profiles/ncu_all_r2.json

It also writes the ALGORITHMIC bytes of every launch (what the kernel must move by definition of its job) to
gpurun_out/prof_all_bytes.json, keyed by kernel name, so that the summary can state dram_bytes / algorithmic_bytes.
Nothing printed under ncu is a benchmark value."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pasture_b200 as pb  # noqa: E402
from pasture_b200 import algorithms as alg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=50_000_000)
    ap.add_argument("--knn-points", type=int, default=4_000_000)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    n, m = args.points, args.knn_points
    want = (lambda k: not args.only or k in args.only.split(","))
    by = {}  # kernel name -> {"algorithmic_bytes": per launch, "workload": ...}

    def note(kernel, nbytes, workload):
        by[kernel] = {"algorithmic_bytes": float(nbytes), "workload": workload}

    raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    sc, of = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    if want("convert"):
        src = alg.synth_las_fmt0_records(n)
        col = pb.HashMapBuffer(tgt, n, "cuda")
        cv = pb.get_default_las_converter(raw, tgt, sc, of)
        cv.convert_into(src, col)  # C2 direction
        note("convert_tiles_kernel#1", 55 * n, "C2: interleaved raw LAS fmt0 (20 B) -> columnar LasPointFormat0 (35 B)")
        mm = torch.zeros(6, dtype=torch.float64, device="cuda")
        cv.convert_into_range_with_bounds_device(src, range(0, n), col, range(0, n), mm)
        note("convert_tiles_kernel#2", 55 * n, "C5 per-GPU step: C2 + fused AABB of the produced POSITION_3D")
        del src
        aos = pb.VectorBuffer(tgt, n, "cuda")
        ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
        ident.convert_into(col, aos)
        note("convert_tiles_kernel#3", 70 * n, "columnar -> interleaved LasPointFormat0 (35 B packed records)")
        back = pb.VectorBuffer(raw, n, "cuda")
        wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
        wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION,
                                                  pb.InvScaleOffset(0.001, of), True)
        wr.convert_into(aos, back)
        note("convert_tiles_kernel#4", 55 * n, "C1 on the GPU: interleaved 35 B -> raw LAS fmt0 20 B, (p-o)/s truncating, convert_into (read-modify-write)")
        wr.convert_into_fresh(aos, back)
        note("convert_tiles_kernel#5", 55 * n, "C1 on the GPU, convert semantics (fresh target: zero ops, no read-modify-write)")
        del aos, back
        from pasture_b200 import las
        las.write_points(col, 0, sc, of)
        note("convert_tiles_kernel#6", 55 * n, "LAS egress: columnar default layout -> fmt0 records, bit-field packing, out-of-range count, points by return and bounds fused")
        del col
    if want("reduce"):
        src = alg.synth_terrain_positions(n)
        alg.calculate_bounds(src)
        note("bounds_flat_f64_kernel", 24 * n, "AABB of a packed Vec3f64 column")
        alg.morton_codes(src, (0.0, 0.0, -10.0), (500.0, 500.0, 16.0))
        note("morton_kernel", 32 * n, "63-bit Morton codes (24 B in, 8 B out)")
        l2 = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.INTENSITY])
        b2 = pb.VectorBuffer(l2, n // 2, "cuda")
        b2.data.view(torch.int16)[:].random_(0, 30000)
        alg.minmax_attribute(b2, pb.attributes.INTENSITY)
        note("minmax_strided_kernel", 2 * (n // 2), "minmax_attribute(INTENSITY u16) on 32 B interleaved records (strided: touches every line)")
        del b2
        alg.voxelgrid_filter(src, 0.1, 0.1, 0.1)
        note("voxel_key_kernel", 32 * n, "C3 voxel keys (24 B in, 8 B out)")
        note("radix_histogram_kernel", 8 * n, "C3 sort: digit histograms of all passes")
        note("radix_onesweep_kernel", 16 * n, "C3 sort: one 8- or 9-bit pass, keys only (8 B in, 8 B out)")
        note("heads_count_kernel", 8 * n, "C3 voxel boundaries: heads per tile")
        note("voxel_emit_reduce_kernel", 32 * n + 32 * 0.4 * n, "C3 boundaries + per-voxel centroid fused (8 B key + 24 B gathered position per point, 32 B per voxel)")
        # a second call with other attributes: the index hand-out path, the thread-per-voxel reductions, heads_emit when
        # the target has no Position3D
        l3 = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.INTENSITY, pb.attributes.CLASSIFICATION])
        m3 = n // 5
        b3 = pb.HashMapBuffer(l3, m3, "cuda")
        b3.columns[0][: 24 * m3].copy_(src.columns[0][: 24 * m3])
        b3.columns[1][: 2 * m3].random_(0, 255)
        b3.columns[2][:m3].random_(0, 5)
        alg.voxelgrid_filter(b3, 0.1, 0.1, 0.1)
        note("voxel_reduce_kernel", (4 + 2) * m3 + 2 * 0.75 * m3, "voxel mean of a u16 attribute (4 B index + 2 B gathered value per point, 2 B per voxel), 10 M points")
        del b3
        del src
    if want("reproject"):
        src = alg.synth_terrain_positions(n // 2)
        p = src.columns[0][: 24 * (n // 2)].view(torch.float64).view(-1, 3)
        p[:, 0].mul_(0.01).add_(35.0)
        p[:, 1].mul_(0.01).sub_(120.0)
        alg.reproject_point_cloud_within(src, "EPSG:4326", "EPSG:3309")
        note("reproject_kernel", 48 * (n // 2), "EPSG:4326 -> EPSG:3309 (geodetic -> ECEF -> shift -> geodetic -> Albers), 24 B in + 24 B out")
        del src, p
    if want("filter"):
        src = alg.synth_las_fmt0_records(n)
        rec = src.data[: 20 * n].view(n, 20)
        mask = ((rec[:, 15] & 1) == 1).to(torch.uint8)
        kept = int(mask.sum().item())
        dst = pb.VectorBuffer(src.point_layout(), kept, "cuda")
        pb.filter_into(src, dst, mask)
        note("filter_count_kernel", 1 * n, "filter_into: kept points per tile (mask read)")
        note("filter_compact_kernel", n + 20 * n + 20 * kept, "filter_into: in-tile compaction of 20 B records, ~50% kept")
        del src, dst, mask, rec
    if want("ransac"):
        src = alg.synth_terrain_positions(m * 5)
        rng = np.random.default_rng(1)
        samples = rng.integers(0, m * 5, (256, 3)).astype(np.uint64)
        alg.ransac_rank_samples(src, 0, samples, 0.5)
        note("ransac_rank_kernel", 24 * m * 5, "RANSAC plane: 256 models ranked in one pass over 20 M positions (FP64-bound, see DESIGN)")
        del src
    if want("pnts"):
        from pasture_b200 import tiles3d
        _l = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.COLOR_RGB])
        src = pb.HashMapBuffer(_l, n, "cuda")
        src.columns[0][: 24 * n].view(torch.float64).uniform_(-1000.0, 1000.0)
        tiles3d.PntsWriter(_l).write(src)
        note("convert_tiles_kernel#7", 45 * n, ".pnts egress: Vec3f64 + Vec3u16 columns -> Vec3f32 + Vec3u8 body")
        del src
    if want("knn"):
        src = alg.synth_terrain_positions(m)
        alg.compute_normals(src, 16)
        note("radix_onesweep_kernel<pairs>", 24 * m, "LBVH sort: one 8-bit pass of (63-bit code, u32 index) pairs")
        note("lbvh_codes_kernel", 36 * m, "LBVH: Morton code + index per point")
        note("gather_positions_kernel", 52 * m, "LBVH: positions gathered into Morton order")
        note("lbvh_hierarchy_kernel", (8 + 64) * m / 8, "LBVH: Karras hierarchy over the bucket codes (one node per 8 points)")
        note("lbvh_refit_kernel", (24 * 8 + 64) * m / 8, "LBVH: bottom-up box refit")
        note("lbvh_query_kernel", 56 * m, "kNN k=16 + normals, packet traversal (compulsory bytes only: latency-bound)")
        del src
    # kernels that are launched several times get one entry per launch ("name#k", k = order of launch in this script)
    m3, vn = n // 5, 0.4
    for k, (nb, w) in enumerate([(24 * n, "AABB of a packed Vec3f64 column"), (24 * n, "AABB (voxel grid call)"), (24 * m3, "AABB, 10 M points"),
                                 (24 * m, "AABB (LBVH build)")], 1):
        note(f"bounds_flat_f64_kernel#{k}", nb, w)
    for k, (nb, w) in enumerate([(32 * n, "C3 voxel keys (24 B in, 8 B out)"), (32 * m3, "voxel keys, 10 M points")], 1):
        note(f"voxel_key_kernel#{k}", nb, w)
    for k, (nb, w) in enumerate([(8 * n, "sort: digit histograms of all passes"), (8 * m3, "10 M keys"), (8 * m, "LBVH codes")], 1):
        note(f"radix_histogram_kernel#{k}", nb, w)
    for k in range(1, 17):
        if k <= 4:
            note(f"radix_onesweep_kernel#{k}", 16 * n, "C3 sort: one pass, keys only (8 B in, 8 B out)")
        elif k <= 8:
            note(f"radix_onesweep_kernel#{k}", 16 * m3, "one pass, keys only, 10 M keys")
        else:
            note(f"radix_onesweep_kernel#{k}", 24 * m, "LBVH sort: one 8-bit pass of (code, index) pairs")
    note("heads_count_kernel#1", 8 * n, "voxel boundaries: heads per tile")
    note("heads_count_kernel#2", 8 * m3, "voxel boundaries, 10 M keys")
    note("voxel_emit_reduce_kernel#1", 32 * n + 32 * vn * n, "C3 boundaries + per-voxel centroid fused (8 B key + 24 B gathered position per point, 32 B per voxel)")
    note("voxel_emit_reduce_kernel#2", 36 * m3 + 36 * 0.75 * m3, "same with index hand-out, 10 M points")
    note("voxel_reduce_kernel#1", 6 * m3 + 2 * 0.75 * m3, "voxel mean of a u16 attribute (4 B index + 2 B gathered value per point)")
    note("voxel_reduce_kernel#2", 5 * m3 + 0.75 * m3, "voxel mode of a u8 attribute (4 B index + 1 B gathered value per point)")
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "prof_all_bytes.json"), "w") as fh:
        json.dump({"points": n, "knn_points": m, "kernels": by}, fh, indent=1)
    print("prof_all done", len(by))


if __name__ == "__main__":
    main()
