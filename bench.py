#!/usr/bin/env python
"""bench.py -- headline benchmark of the pasture hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): 100 M raw LAS format-0 records (interleaved, 20 B)
-> columnar LasPointFormat0 attributes (10 columns, 35 B/point) with the default LAS read mappings
(i32 -> f64 cast then v*scale+offset on POSITION_3D, four bit-field extracts, five copies), i.e.
pasture's BufferLayoutConverter::convert_into_range as configured by get_default_las_converter
(pasture-io/src/las/raw_readers.rs:31-167).  Algorithmic bytes: 20 read + 35 written = 55 B/point.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            the reference algorithm on the host cores (rank 0 only)

N > 1: every rank converts its own 100 M-point shard (weak scaling, shard r = point indices
r*1e8 ...), the AABB of the produced positions is fused into the convert kernel and one 48-byte NCCL
all-reduce(min) of [min xyz, -max xyz] gives the global bounds.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

POINTS_PER_GPU = 100_000_000
BYTES_IN, BYTES_OUT = 20, 35
SCALE = (0.001, 0.001, 0.001)
OFFSET = (500000.0, 5400000.0, 100.0)
METRIC = "points/sec layout-convert+transform (interleaved LAS fmt0 -> columnar, scale/offset)"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes(fused):
    """dram bytes per launch of the convert kernel from the committed `ncu --set full` capture of the same workload (a
    static number read from profiles/, not measured in this run: ncu cannot run inside a timed bench).  C2 and C5 (fused
    AABB) have separate captures; returns (bytes or None, source label)"""
    name = "convert_c5_ncu_summary.json" if fused else "convert_c2_ncu_summary.json"
    p = os.path.join(ROOT, "profiles", name)
    try:
        with open(p) as fh:
            return float(json.load(fh)["dram_bytes_per_launch"]), f"profiles/{name} (static: committed ncu --set full capture)"
    except Exception:
        return None, f"none (no committed capture profiles/{name})"


class ClockSampler:
    """samples SM clock / throttle reasons during the timed region (pynvml, else nvidia-smi)"""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        if nv is not None:
            for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown",
                      "SyncBoost", "ApplicationsClocksSetting", "DisplayClockSetting"):
                v = getattr(nv, "nvmlClocksThrottleReason" + n, None)
                if v is not None:
                    names[int(v)] = n
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for bit, n in names.items():
                        if r & bit:
                            self.reasons.add(n)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for name, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002 if nv is not None else 0.05)  # the timed region is ~20 ms: NVML is polled every 2 ms

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_row_rates():
    """The same CPU port timed on the other SURVEY 8a rows (single thread, bounded samples, a few seconds in total), so
    that the GPU numbers of benchmarks/bench_configs.py (profiles/configs_r1.jsonl) have the reference algorithm's CPU
    cost beside them: AABB (bounds.rs), voxel grid as written (quadratic insert, voxel_grid.rs) and as a sort-based
    restatement, kNN normals by brute force (normal_estimation.rs with an exact kNN)."""
    import numpy as np
    import oracle as O
    rows = {}

    def rate(n, fn):
        t0 = time.perf_counter()
        fn()
        return {"points": n, "seconds": round(time.perf_counter() - t0, 3), "points_per_s": n / max(time.perf_counter() - t0, 1e-9)}

    ol = O.OLayout.from_attributes([("Position3D", O.VEC3F64)])

    def cloud(n):
        b = O.OBuffer(ol, n, True)
        b.set_attribute("Position3D", O.gen_terrain_positions(0, n))
        return b

    b = cloud(5_000_000)
    rows["aabb (calculate_bounds, C3 stream)"] = rate(5_000_000, lambda: O.calculate_bounds(b))
    b2 = cloud(2_000_000)
    rows["voxel grid 0.1 m, sort-based restatement"] = rate(2_000_000, lambda: O.voxelgrid_filter(b2, (0.1, 0.1, 0.1), ol, use_sort=True))
    b3 = cloud(20_000)
    rows["voxel grid 0.1 m, as written in the reference (sorted-vector insert per point)"] = rate(
        20_000, lambda: O.voxelgrid_filter(b3, (0.1, 0.1, 0.1), ol))
    pts = O.gen_terrain_positions(0, 8_000)
    rows["normals k=16 with brute-force exact kNN"] = rate(8_000, lambda: O.compute_normals(pts, 16))
    return rows


# algorithmic bytes per point of every timed phase (SURVEY 8d conventions; DESIGN.md section 4 states them per kernel).
# V/N = voxels per point enters the C3 phases that write per-voxel results.
def _c3_phase_bytes(v_over_n, passes):
    return {"voxel.bounds": 24, "voxel.keys": 24 + 8, "sort.histogram": 8, "sort.pass": 16,
            "voxel.heads_count+scan": 8, "voxel.heads_emit": 8 + 4 + 12 * v_over_n,
            "voxel.reduce": 4 + 24 + (24 + 4) * v_over_n, "voxel.emit+reduce": 8 + 24 + (24 + 8) * v_over_n}


_C4_PHASE_BYTES = {"knn.bounds": 24, "knn.codes": 24 + 12, "sort.histogram": 8, "sort.pass": 24,
                   "knn.gather_positions": 4 + 24 + 24, "knn.tree": 8 / 8 + 2 * 64 / 8, "knn.query+normals": 24 + 32}


def _phase_table(phases, bytes_of, n, peak):
    """[(name, ms)] of ONE call -> per-kernel rows; repeated names (sort passes) are summed"""
    agg = {}
    phases = [(nm, ms) for nm, ms in phases if not nm.endswith(".call") and not nm.startswith(("host.", "pool."))]  # kernels only
    for name, ms in phases:
        a = agg.setdefault(name, {"name": name, "ms": 0.0, "launch_groups": 0})
        a["ms"] += ms
        a["launch_groups"] += 1
    rows = []
    for name, a in agg.items():
        b = bytes_of.get(name)
        if b is not None:
            a["algorithmic_bytes"] = b * n * a["launch_groups"]
            a["achieved_GBps"] = a["algorithmic_bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else None
            a["frac"] = a["achieved_GBps"] / peak if a["achieved_GBps"] else None
        rows.append(a)
    return rows


def other_configs(pb, ctx, dev, peak):
    """C3 (100 M-point AABB + voxel keys + sort + voxel-grid downsample at 0.1 m, voxel_grid.rs:109-165) and C4 (100 M-point
    kNN k=16 normal estimation over the device LBVH, normal_estimation.rs:79-130) measured after the headline, each with
    its own clock sample and the per-phase CUDA-event times of the library's phase timer (pb200_ctx_profile_read)."""
    import torch
    from pasture_b200 import algorithms as alg
    out = []
    n = POINTS_PER_GPU

    def run(label, fn, bytes_total_per_point, phase_bytes, reps):
        fn()  # warm-up: first call allocates the temporaries from the driver
        torch.cuda.synchronize()
        sampler = ClockSampler(dev.index)
        times = []
        with sampler as clocks:
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = fn()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        ctx.profile(True)
        res = fn()
        phases = ctx.profile_read()
        ctx.profile(False)
        ms = min(times)
        row = {"workload": label, "points": n, "ms": ms, "ms_all": times, "points_per_s": n / (ms * 1e-3),
               "algorithmic_bytes_per_point": bytes_total_per_point,
               "achieved_GBps": bytes_total_per_point * n / (ms * 1e-3) / 1e9,
               "frac": bytes_total_per_point * n / (ms * 1e-3) / 1e9 / peak,
               "kernels": _phase_table(phases, phase_bytes(res) if callable(phase_bytes) else phase_bytes, n, peak),
               "kernel_ms_sum": sum(m for nm, m in phases if not nm.endswith(".call") and not nm.startswith(("host.", "pool."))),
               "pool_MB": {nm: m for nm, m in phases if nm.startswith("pool.")},
               "call_ms_profiled": next((m for nm, m in phases if nm.endswith(".call")), None),
               "host_timeline_us": [(nm, round(a), round(b)) for nm, a, b in ctx.last_profile_host_us if not nm.startswith("pool.")],
               "clocks": clocks.summary(),
               "timing": "CUDA events around the whole library call (best of %d, device-resident input, includes the "
                         "call's host synchronisations); kernels: one extra profiled call" % reps}
        return row, res

    src = alg.synth_terrain_positions(n, device=dev)
    row, res = run("C3: 100M-point AABB + voxel keys + radix sort + voxel-grid downsample (0.1 m) on 1xB200",
                   lambda: alg.voxelgrid_filter(src, 0.1, 0.1, 0.1), 232,
                   lambda r: _c3_phase_bytes(r.len() / n, 5), reps=5)
    row["voxels"] = res.len()
    out.append(row)
    del res
    ctx.trim()
    row, res = run("C4: 100M-point kNN (k=16) normal estimation over the device LBVH on 1xB200",
                   lambda: alg.compute_normals(src, 16), 56, _C4_PHASE_BYTES, reps=2)
    row["note"] = ("not HBM-bound (tree traversal is latency / L1-bound): frac is informational (SURVEY 8d); the LBVH build "
                   "phases (codes, sort, gather, tree) are HBM-class")
    out.append(row)
    del res, src
    ctx.trim()
    out.append(converter_directions(pb, dev, peak))
    ctx.trim()
    return out


def converter_directions(pb, dev, peak):
    """the other loops of BufferLayoutConverter (buffer_conversion.rs:418-662) and the two egress paths at 100 M points, default
    schedule: columnar -> interleaved packed records, interleaved packed records -> columnar, the write direction (C1 on the
    GPU) with `convert` and `convert_into` semantics, LAS egress, .pnts egress.  One clock sample over the block."""
    import torch
    from pasture_b200 import algorithms as alg, las, tiles3d
    n = POINTS_PER_GPU
    raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    rows = []

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        return best

    def emit(name, ms, bpp, note=None):
        r = {"direction": name, "ms": ms, "algorithmic_bytes_per_point": bpp, "achieved_GBps": bpp * n / (ms * 1e-3) / 1e9,
             "frac": bpp * n / (ms * 1e-3) / 1e9 / peak}
        if note:
            r["note"] = note
        rows.append(r)

    with ClockSampler(dev.index) as clocks:
        src = alg.synth_las_fmt0_records(n, device=dev)
        col = pb.HashMapBuffer(tgt, n, dev)
        pb.get_default_las_converter(raw, tgt, SCALE, OFFSET).convert_into(src, col)
        del src
        aos = pb.VectorBuffer(tgt, n, dev)
        ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
        emit("columnar -> interleaved LasPointFormat0 (35 B packed records)", timed(lambda: ident.convert_into(col, aos)), 70)
        col2 = pb.HashMapBuffer(tgt, n, dev)
        emit("interleaved LasPointFormat0 (35 B packed records) -> columnar", timed(lambda: ident.convert_into(aos, col2)), 70)
        del col2
        back = pb.VectorBuffer(raw, n, dev)
        wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
        wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION,
                                                  pb.InvScaleOffset(SCALE[0], OFFSET), True)
        emit("C1 on the GPU: interleaved 35 B -> raw LAS fmt0 20 B, (p-o)/s truncating, `convert` semantics (fresh target)",
             timed(lambda: wr.convert_into_fresh(aos, back)), 55)
        emit("same, `convert_into` semantics (unmapped target bytes preserved)", timed(lambda: wr.convert_into(aos, back)), 55,
             "read-modify-write of the partially mapped records: 75 B/pt of real traffic")
        del aos, back
        emit("LAS egress: columnar default layout -> fmt0 records + out-of-range count + points by return + bounds",
             timed(lambda: las.write_points(col, 0, SCALE, OFFSET), reps=3), 55)
        del col
        lay = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.COLOR_RGB])
        s2 = pb.HashMapBuffer(lay, n, dev)
        s2.columns[0][: 24 * n].view(torch.float64).uniform_(-1000.0, 1000.0)
        w = tiles3d.PntsWriter(lay)

        def pnts():
            w._chunks.clear()
            w.write(s2)
        emit(".pnts egress: Vec3f64 + Vec3u16 columns -> FeatureTable body (Vec3f32 + Vec3u8)", timed(pnts, reps=3), 45)
        del s2, w
    return {"workload": "converter directions and egress paths at 100M points on 1xB200 (default schedule)", "points": n,
            "rows": rows, "clocks": clocks.summary(),
            "timing": "CUDA events around the library call, best of 5 (egress: 3) after one warm-up, device-resident buffers"}


def oracle_convert_rate(n_points, threads, repeats=1):
    """times the CPU restatement of convert_into_range (attribute-outer / point-inner, function-pointer casts,
    buffer_conversion.rs:546-604) on n_points of the C2 stream; returns (points/s, seconds of the best repeat)"""
    import oracle as O
    ol_raw, ol_def = O.OLayout.las_raw(0), O.OLayout.las_default(0)
    src = O.OBuffer(ol_raw, n_points, False)
    src.aos[:] = O.gen_las_fmt0_records(0, n_points, 42)
    dst = O.OBuffer(ol_def, n_points, True)
    cv = O.OConverter.las_default(ol_raw, ol_def, SCALE, OFFSET)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cv.convert_into_range(src, 0, n_points, dst, 0, n_points, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return n_points / best, best


def run_reference(args):
    """--impl reference: the reference's algorithm for this path on the host cores. The Rust reference cannot be
    built in this image (no rustc/cargo), so this is the oracle port (kind "port"): the same per-range routine
    over all host threads (what wrapping convert_into_range in rayon would give; the reference itself is
    single-threaded for this path)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 20_000_000
    import oracle as O
    ol_raw, ol_def = O.OLayout.las_raw(0), O.OLayout.las_default(0)
    src = O.OBuffer(ol_raw, sample, False)
    src.aos[:] = O.gen_las_fmt0_records(0, sample, 42)
    dst = O.OBuffer(ol_def, sample, True)
    cv = O.OConverter.las_default(ol_raw, ol_def, SCALE, OFFSET)
    for _ in range(max(1, args.warmup)):
        cv.convert_into_range(src, 0, sample, dst, 0, sample, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cv.convert_into_range(src, 0, sample, dst, 0, sample, threads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} points of the C2 stream per step, {cores} threads over disjoint point ranges, host memory"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 100M-point interleaved->columnar convert + scale/offset (bounded CPU sample)",
                   "points_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import pasture_b200 as pb
    from pasture_b200.algorithms import synth_las_fmt0_records

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    ctx = pb.get_context(dev.index)
    # pinned host buffers of the e2e leg should live on the GPU's NUMA node: bind this rank's thread to the device-local
    # CPUs before anything is allocated (no effect on single-node VMs; recorded in e2e.host_binding)
    host_binding = None if args.no_bind else ctx.bind_host_thread()
    for kv in args.param:
        k, v = kv.split("=")
        ctx.set_param(k, int(v))
    n = args.points
    fused = world > 1 or args.fused_bounds

    pl_raw, pl_def = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    src = synth_las_fmt0_records(n, first_index=rank * n, seed=42, device=dev)  # resident in HBM before timing
    dst = pb.HashMapBuffer(pl_def, n, dev)
    cv = pb.get_default_las_converter(pl_raw, pl_def, SCALE, OFFSET)
    minmax6 = torch.zeros(6, dtype=torch.float64, device=dev)
    rng_all = range(0, n)

    # N > 1: the global AABB is exchanged over peer memory by the last CTA of the convert kernel itself (PeerComm: CUDA-IPC
    # mapped exchange buffers, NVLink stores) -- no collective-library call inside a step.  --nccl-bounds (or a failed IPC
    # set-up, agreed on by all ranks) falls back to the fused kernel + one 48-byte NCCL all-reduce(MIN).
    comm, exchange = None, "none (single GPU)"
    if world > 1:
        exchange = "NCCL all-reduce(MIN), 48 B"
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        if not args.nccl_bounds:
            try:
                from pasture_b200.sharding import PeerComm
                comm = PeerComm(ctx)
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write(f"[rank {rank}] peer-memory communicator unavailable ({exc}); using NCCL\n")
                ok.zero_()
                comm = None
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                exchange = "peer memory: last CTA of the convert kernel stores its 6 keys into every peer's buffer over NVLink (no NCCL call per step)"
            else:
                comm = None

    def step():
        if comm is not None:
            cv.convert_into_range_with_global_bounds(src, rng_all, dst, rng_all, comm, minmax6)
        elif fused:
            cv.convert_into_range_with_bounds_device(src, rng_all, dst, rng_all, minmax6)
            if world > 1:
                dist.all_reduce(minmax6, op=dist.ReduceOp.MIN)
        else:
            cv.convert_into_range(src, rng_all, dst, rng_all)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching stream (torch's current stream) ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = pb.kernel_launch_count()
    sampler = ClockSampler(dev.index)  # NVML initialisation takes milliseconds and differs per rank: do it BEFORE the barrier
    with sampler as clocks:
        if world > 1:  # all ranks enter the timed region together: at N > 1 every step waits for the slowest rank, so a
            dist.barrier()  # start skew of a few ms would be charged to every rank's K steps
            torch.cuda.synchronize()
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    launches = pb.kernel_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- sanity: the produced bounds must be the bounds of the shard (cheap device-side check) ----------
    if comm is not None:
        comm.check()  # raises if a peer never arrived inside the kernel's timeout
    if fused:
        mm = minmax6.cpu().numpy()
        assert mm[0] <= -mm[3] and mm[1] <= -mm[4] and mm[2] <= -mm[5], mm
        if world > 1:  # every rank must hold the same, global box; rank 0 also checks it against its own shard's box
            allmm = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allmm, minmax6)
            assert all(bool((a == allmm[0]).all()) for a in allmm), "ranks disagree on the global AABB"
            local6 = torch.zeros(6, dtype=torch.float64, device=dev)
            cv.convert_into_range_with_bounds_device(src, rng_all, dst, rng_all, local6)
            ref6 = local6.clone()
            dist.all_reduce(ref6, op=dist.ReduceOp.MIN)
            assert bool((ref6 == minmax6).all()), (ref6, minmax6)

    # ---- roofline of the dominant kernel (the only kernel in a step at N = 1) -----------------------------
    peak, peak_src = measured_peak_gbs()
    avg_kernel_ms = sum(step_ms) / len(step_ms)
    best_kernel_ms = min(step_ms)
    achieved = (BYTES_IN + BYTES_OUT) * n / (avg_kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_bytes(fused)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": (BYTES_IN + BYTES_OUT) * n,
                "avg_launch_ms": avg_kernel_ms, "best_launch_ms": best_kernel_ms,
                "frac_of_8TBps_nominal": achieved / 8000.0, "kernel": "convert_tiles_kernel",
                "note": "rank 0; at N>1 the launch also holds the fused AABB and its exchange" if world > 1 else "rank 0"}

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H every step ------------
    # every rank runs it at the same time (all GPUs share the host's memory system), time = max over ranks.
    # Right before it, benchmarks/pcie_probe.py measures what plain pinned cudaMemcpyAsync of the same byte counts achieves
    # in the same process layout (all ranks concurrently): its both-directions time is the floor of an e2e step, and
    # e2e.roofline.frac = floor / measured says how much of the remaining time is the library's.
    e2e = None
    if not args.no_e2e:
        n_e2e = args.e2e_points
        probe = None
        if not args.no_pcie_probe:
            from benchmarks.pcie_probe import probe as pcie_probe
            probe = pcie_probe(dev, n_e2e * BYTES_IN, n_e2e * BYTES_OUT, reps=3, dist=dist if world > 1 else None)
        h_src = pb.VectorBuffer(pl_raw, n_e2e, "cpu", pinned=True)
        h_src.data[: n_e2e * BYTES_IN].copy_(src.data[: n_e2e * BYTES_IN])
        h_dst = pb.HashMapBuffer(pl_def, n_e2e, "cpu", pinned=True)
        torch.cuda.synchronize()
        r = range(0, n_e2e)
        cv.convert_into_range(h_src, r, h_dst, r)  # warm-up (allocates the staging buffers)
        e2e_steps = max(1, min(args.steps, 10))
        per_step = []
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            cv.convert_into_range(h_src, r, h_dst, r)  # returns after the D2H of the last chunk
            per_step.append((time.perf_counter() - ts) * 1e3)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        per_step.sort()
        e2e = {"value": world * n_e2e * e2e_steps / dt, "unit": "points/s", "h2d_bytes_per_step": n_e2e * BYTES_IN,
               "d2h_bytes_per_step": n_e2e * BYTES_OUT, "ms_per_step": dt / e2e_steps * 1e3, "points_per_step": n_e2e,
               "steps": e2e_steps,
               "step_ms_rank0": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
               "host_binding": host_binding,
               "note": "per rank: pinned host VectorBuffer -> pb200_converter_convert_into_range (HOST memspace, chunked "
                       "H2D/kernel/D2H overlap) -> pinned host HashMapBuffer; all ranks concurrently, max over ranks"}
        if probe is not None:
            floor_ms = probe["both_ms"]
            e2e["roofline"] = {
                "bound": "pcie/host-memory", "floor_ms": floor_ms, "frac": floor_ms / (dt / e2e_steps * 1e3),
                "peak_gbs_h2d": probe["h2d_alone_gbs"], "peak_gbs_d2h": probe["d2h_alone_gbs"],
                "concurrent_gbs_h2d": probe["both_h2d_gbs"], "concurrent_gbs_d2h": probe["both_d2h_gbs"],
                "solo_floor_ms": probe.get("solo_both_ms"),
                "source": "benchmarks/pcie_probe.py run in this process right before the e2e leg: plain pinned "
                          "cudaMemcpyAsync of the same H2D and D2H byte counts on two streams, all ranks concurrently, "
                          "per-rank GB/s, max time over ranks (solo = rank 0 alone)"}
        # the host result must equal the device result
        for i in (0, len(pl_def) - 1):
            a = h_dst.columns[i][: n_e2e * pl_def.at(i).size()]
            b = dst.columns[i][: n_e2e * pl_def.at(i).size()].cpu()
            assert torch.equal(a, b), f"e2e column {i} differs from the device-resident result"
        del h_src, h_dst

    # ---- N > 1: what the line needs to be read on its own ---------------------------------------------------------
    # per-rank step times (which rank the others wait for) and, on every rank, the single-GPU cost of the fused AABB
    # (C2 with and without min/max tracking, no exchange): the N = 1 -> N >= 2 step is that, not communication.
    attribution = None
    if world > 1:
        def local_ms(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        plain_ms = local_ms(lambda: cv.convert_into_range(src, rng_all, dst, rng_all))
        fused_ms = local_ms(lambda: cv.convert_into_range_with_bounds_device(src, rng_all, dst, rng_all, minmax6))
        ss = sorted(step_ms)
        mine = {"rank": rank, "step_ms": {"min": ss[0], "median": ss[len(ss) // 2], "max": ss[-1]},
                "c2_plain_ms": plain_ms, "c2_fused_aabb_ms": fused_ms}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        attribution = {"per_rank": allr, "fused_aabb_ms": allr[0]["c2_fused_aabb_ms"], "plain_ms": allr[0]["c2_plain_ms"],
                       "note": "c2_plain_ms / c2_fused_aabb_ms: the same conversion on this rank alone without / with the fused "
                               "AABB (no exchange, not coupled to the peers), 20 launches after the timed region"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the other single-GPU configurations of BASELINE.json at full size (outside the headline's timed region) ----
    other = None
    if world == 1 and not args.no_other_configs:
        del src, dst
        torch.cuda.empty_cache()
        other = other_configs(pb, ctx, dev, peak)

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate1, secs1 = oracle_convert_rate(2_000_000, 1)
        sample = int(min(50_000_000, max(2_000_000, rate1 * 12)))  # ~12 s of single-thread work
        rate, secs = oracle_convert_rate(sample, 1)
        cores = os.cpu_count() or 1
        mt_sample = int(min(50_000_000, max(sample, rate * cores * 0.5)))
        rate_mt, secs_mt = oracle_convert_rate(mt_sample, cores)
        cpu = {"value": rate, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": f"{sample} points of the C2 stream, single thread (the reference path is single-threaded), {secs:.1f} s",
               "mt_value": rate_mt, "mt_cores": cores,
               "mt_sample": f"{mt_sample} points, {cores} threads over disjoint point ranges, {secs_mt:.1f} s",
               "other_rows": oracle_row_rates()}

    out = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("C5: per-GPU 100M-point shard convert + fused AABB + global AABB exchange" if world > 1 else
                                "C2: 100M-point interleaved->columnar convert + scale/offset on 1xB200" +
                                (" + fused AABB" if fused else "")),
                   "points_per_gpu": n, "bounds_exchange": exchange, "source": "VectorBuffer raw LAS fmt0 (20 B/pt)",
                   "target": "HashMapBuffer LasPointFormat0 (10 columns, 35 B/pt)", "parallelism": f"point-range shards x{world}",
                   "l2_policy": "inputs (2.0 GB) and outputs (3.5 GB) per step are far larger than the 126 MB L2"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if attribution is not None:
        out["scaling_attribution"] = attribution
    if other is not None:
        out["other_configs"] = other
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=POINTS_PER_GPU, help="points per GPU (default: the 100 M of C2)")
    ap.add_argument("--e2e-points", type=int, default=POINTS_PER_GPU)
    ap.add_argument("--fused-bounds", action="store_true", help="N=1: also fuse the AABB (always on for N>1)")
    ap.add_argument("--nccl-bounds", action="store_true", help="N>1: exchange the AABB with an NCCL all-reduce instead of peer memory")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pcie-probe", action="store_true", help="skip the pinned-copy probe that gives e2e.roofline")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to its GPU's NUMA-local CPUs")
    ap.add_argument("--no-other-configs", action="store_true", help="N=1: skip the C3 / C4 legs (other_configs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--param", action="append", default=[], help="tuning knob key=value (pb200_ctx_set_param), e.g. convert.threads=512")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
