#!/usr/bin/env python
"""Informational timings of the non-headline configs of BASELINE.json (C3: AABB + voxel-grid downsample, C4: kNN
normals, plus the stand-alone reductions and the other conversion directions). Prints one JSON object per line;
the headline metric lives in bench.py."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pasture_b200 as pb  # noqa: E402
from pasture_b200 import algorithms as alg  # noqa: E402


LAST_CLOCKS = None


def timed(fn, reps=3, warm=1):
    """best of `reps` CUDA-event timings; the SM clock / throttle reasons are sampled over the timed repetitions (NVML,
    bench.py's ClockSampler) and attached to the next emitted line"""
    global LAST_CLOCKS
    from bench import ClockSampler
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    with ClockSampler(torch.cuda.current_device()) as clocks:
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
    LAST_CLOCKS = clocks.summary()
    return best


def fp64_peak_ginstr():
    """measured FP64 instruction rate (benchmarks/fp64_probe.cu, non-fused DMUL/DADD), G instructions/s; None if the probe
    binary is missing"""
    import subprocess
    exe = os.path.join(ROOT, "benchmarks", "build", "fp64_probe")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        return {r["variant"]: r["G_fp64_instr_per_s"] for r in rows}
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--knn-points", type=int, default=20_000_000)
    ap.add_argument("--skip", default="")
    ap.add_argument("--params", default="", help="context parameters for tuning sweeps, e.g. convert.threads=1024,convert.ctas_per_sm=2")
    args = ap.parse_args()
    for kv in filter(None, args.params.split(",")):
        k, v = kv.split("=")
        pb.get_context().set_param(k, int(v))
    n = args.points
    peak = 6536.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def tuned(fn, reps=3):
        """the same call with the library's schedule autotuning on ("convert.autotune": the first large conversion of a plan
        shape times alternative schedules on a prefix of its own range and the context keeps the fastest)"""
        ctx = pb.get_context()
        ctx.set_param("convert.autotune", 1)
        try:
            fn()  # tunes
            return timed(fn, reps=reps)
        finally:
            ctx.set_param("convert.autotune", 0)

    def emit(name, ms, npts, bytes_per_pt, extra=None):
        d = {"config": name, "points": npts, "ms": ms, "points_per_s": npts / (ms * 1e-3),
             "algorithmic_bytes_per_point": bytes_per_pt, "achieved_GBps": bytes_per_pt * npts / (ms * 1e-3) / 1e9 if bytes_per_pt else None,
             "frac_of_measured_peak": (bytes_per_pt * npts / (ms * 1e-3) / 1e9 / peak) if bytes_per_pt else None}
        if extra:
            d.update(extra)
        d["clocks"] = LAST_CLOCKS
        print(json.dumps(d), flush=True)

    if "aabb" not in args.skip:
        src = alg.synth_terrain_positions(n)
        ms = timed(lambda: alg.calculate_bounds(src))
        emit("AABB of packed Vec3f64 column (calculate_bounds, incl. the 48 B readback)", ms, n, 24)
        ms = timed(lambda: alg.morton_codes(src, (0.0, 0.0, -10.0), (500.0, 500.0, 16.0)))
        emit("63-bit Morton codes of a Vec3f64 column (24 B in, 8 B out; incl. the torch allocation of the output)", ms, n, 32)
        ms = timed(lambda: alg.minmax_attribute(src, pb.attributes.POSITION_3D))
        emit("minmax_attribute(POSITION_3D Vec3f64) on the packed column", ms, n, 24)
        l2 = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.INTENSITY])
        b2 = pb.VectorBuffer(l2, n // 2, "cuda")
        ms = timed(lambda: alg.minmax_attribute(b2, pb.attributes.INTENSITY))
        emit("minmax_attribute(INTENSITY u16) on 32 B interleaved records: strided, every sector of the buffer is touched "
             "(32 B/pt of DRAM traffic; against the 2 B/pt of data it needs the fraction is meaningless, against the touched bytes it is the second figure)",
             ms, n // 2, 2, {"frac_of_measured_peak_on_touched_bytes": 32 * (n // 2) / (ms * 1e-3) / 1e9 / peak})
        del b2
        p = src.columns[0][: 24 * n].view(torch.float64).view(-1, 3)
        p[:, 0].mul_(0.01).add_(35.0)
        p[:, 1].mul_(0.01).sub_(120.0)
        ms = timed(lambda: alg.reproject_point_cloud_within(src, "EPSG:4326", "EPSG:32610"), reps=2)
        emit("reprojection EPSG:4326 -> UTM zone 10N in place (Krueger series; 24 B in + 24 B out; FP64 transcendental-bound)", ms, n, 48)
        del src, p
    if "c3" not in args.skip:
        src = alg.synth_terrain_positions(n)
        t0 = time.perf_counter()
        out = alg.voxelgrid_filter(src, 0.1, 0.1, 0.1)
        torch.cuda.synchronize()
        first = (time.perf_counter() - t0) * 1e3
        ms = timed(lambda: alg.voxelgrid_filter(src, 0.1, 0.1, 0.1), reps=2, warm=0)
        emit("C3: AABB + voxel keys + radix sort + per-voxel centroid, leaf 0.1 m (wall-clock incl. allocations)", ms, n, 232,
             {"voxels": out.len(), "first_call_ms": first})
        del src, out
    if "soa2aos" not in args.skip:
        raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
        src = alg.synth_las_fmt0_records(n)
        col = pb.HashMapBuffer(tgt, n, "cuda")
        cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
        cv.convert_into(src, col)
        del src
        aos = pb.VectorBuffer(tgt, n, "cuda")
        ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
        ms = timed(lambda: ident.convert_into(col, aos))
        emit("columnar -> interleaved (LasPointFormat0 35 B, identity mappings)", ms, n, 70)
        emit("columnar -> interleaved (LasPointFormat0 35 B, identity mappings), schedule autotuned", tuned(lambda: ident.convert_into(col, aos)), n, 70, {"autotune": True})
        col2 = pb.HashMapBuffer(tgt, n, "cuda")
        ms = timed(lambda: ident.convert_into(aos, col2))
        emit("interleaved -> columnar (LasPointFormat0 35 B packed records, identity mappings)", ms, n, 70)
        emit("interleaved -> columnar (35 B packed records), schedule autotuned", tuned(lambda: ident.convert_into(aos, col2)), n, 70, {"autotune": True})
        del col2
        # write direction (C1 on the GPU): 35 B default layout -> 20 B raw records, (p-o)/s truncation
        back = pb.VectorBuffer(raw, n, "cuda")
        wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
        wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION,
                                                  pb.InvScaleOffset(0.001, (500000.0, 5400000.0, 100.0)), True)
        ms = timed(lambda: wr.convert_into_fresh(aos, back))
        emit("C1 on GPU: interleaved LasPointFormat0 (35 B) -> raw LAS fmt0 (20 B), (p-o)/s truncating; `convert` semantics "
             "(fresh target: unmapped flags byte written as 0, no read-modify-write)", ms, n, 55)
        emit("C1 on GPU, `convert` semantics, schedule autotuned", tuned(lambda: wr.convert_into_fresh(aos, back)), n, 55, {"autotune": True})
        ms = timed(lambda: wr.convert_into(aos, back))
        emit("same with `convert_into` semantics (unmapped target bytes preserved: read-modify-write, 75 B/pt of real traffic)", ms, n, 55)
        del col, aos, back
    if "filter" not in args.skip:
        src = alg.synth_las_fmt0_records(n)
        rec = src.data[: 20 * n].view(n, 20)
        mask = ((rec[:, 15] & 1) == 1).to(torch.uint8)
        kept = int(mask.sum().item())
        dst = pb.VectorBuffer(src.point_layout(), kept, "cuda")
        ms = timed(lambda: pb.filter_into(src, dst, mask))
        emit("filter_into: raw LAS fmt0 records (20 B, interleaved), ~50% kept (mask scan + gather)", ms, n, 1 + 20 + 20 * kept / n,
             {"kept": kept})
        del src, dst, mask, rec
    if "ransac" not in args.skip:
        import numpy as np
        m = args.knn_points
        src = alg.synth_terrain_positions(m)
        rng = np.random.default_rng(1)
        fp64 = fp64_peak_ginstr()
        for kind, name in ((0, "plane"), (1, "line")):
            samples = rng.integers(0, m, (300, 3 if kind == 0 else 2)).astype(np.uint64)
            ms = timed(lambda: alg.ransac_rank_samples(src, kind, samples, 0.5), reps=2)
            tests = 300 * m / (ms * 1e-3)
            # FP64-pipe instructions per point-model test of the guard-band path (segmentation.rs:31-44 with every product and
            # sum rounded separately): plane 3 DMUL + 3 DADD + 2 DSETP = 8, line 3 DSUB + 9 DMUL + 3 DSUB + 2 DADD + 2 DSETP = 19
            per_test = 8 if kind == 0 else 19
            extra = {"point_model_tests_per_s": tests, "fp64_instr_per_test": per_test, "fp64_G_instr_per_s": tests * per_test / 1e9,
                     "roofline": "FP64 pipe (HBM fraction is meaningless here: each 24 B position is reused for up to 256 models)"}
            if fp64:
                peak_i = fp64.get("DMUL + DADD (non-fused), 8 chains/thread")
                extra["fp64_peak_G_instr_per_s_measured"] = peak_i
                extra["frac_of_measured_fp64_rate"] = tests * per_test / 1e9 / peak_i if peak_i else None
            emit(f"RANSAC {name}: 300 models ranked over the cloud (2 launches of <=256 models, 24 B/pt each)", ms, m, 48, extra)
        del src
    if "las" not in args.skip:
        from pasture_b200 import las
        raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
        src = alg.synth_las_fmt0_records(n)
        col = pb.HashMapBuffer(tgt, n, "cuda")
        cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
        cv.convert_into(src, col)
        del src
        ms = timed(lambda: las.write_points(col, 0, (0.001,) * 3, (500000.0, 5400000.0, 100.0)), reps=2)
        emit("LAS egress: columnar default layout (35 B) -> fmt0 records (20 B) + counts by return + bounds", ms, n, 55)
        emit("LAS egress, schedule autotuned", tuned(lambda: las.write_points(col, 0, (0.001,) * 3, (500000.0, 5400000.0, 100.0)), reps=2), n, 55, {"autotune": True})
        del col
    if "pnts" not in args.skip:
        from pasture_b200 import tiles3d
        _l = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.COLOR_RGB])
        src = pb.HashMapBuffer(_l, n, "cuda")
        src.columns[0][: 24 * n].view(torch.float64).uniform_(-1000.0, 1000.0)
        w = tiles3d.PntsWriter(_l)
        def wr():
            w._chunks.clear()
            w.write(src)
        ms = timed(wr, reps=2)
        emit(".pnts egress: Vec3f64 + Vec3u16 columns -> FeatureTable body (Vec3f32 + Vec3u8)", ms, n, 30 + 15)
        emit(".pnts egress, schedule autotuned", tuned(wr, reps=2), n, 30 + 15, {"autotune": True})
        del src
    if "c4" not in args.skip:
        m = args.knn_points
        src = alg.synth_terrain_positions(m)
        torch.cuda.empty_cache()  # (the earlier rows' buffers: the first call below allocates its temporaries from the driver)
        t0 = time.perf_counter()
        normals, curv = alg.compute_normals(src, 16)
        torch.cuda.synchronize()
        first = (time.perf_counter() - t0) * 1e3
        del normals, curv
        ms = timed(lambda: alg.compute_normals(src, 16), reps=2, warm=0)
        emit("C4: LBVH build + kNN (k=16) + normals/curvature (whole call; the 100 M-point figure is in bench.py's other_configs)", ms, m, 56,
             {"first_call_ms": first})


if __name__ == "__main__":
    main()
