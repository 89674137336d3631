#!/usr/bin/env python
"""Per-warp timeline of the converter's tile pipeline (diagnostics build only).

The library built with `make -C pasture_b200/csrc TILE_TRACE=1` (use a scratch copy of csrc: the flag changes the kernel's
parameter block) stamps clock64 in CTA 0 for its first 32 tiles: per warp, the time it waited for the tile's data, the time
its work items took, and for thread 0 the time spent issuing the tile's bulk copies.  With PB200_TILE_TRACE=<file> in the
environment every large conversion appends its plan (the items of every warp) and the stamps to <file>.

  python benchmarks/tile_trace.py run   gpurun_out/tile_trace.txt     # on the GPU box, with the diagnostics library in place
  python benchmarks/tile_trace.py show  gpurun_out/tile_trace.txt     # anywhere: markdown summary (medians over tiles 8..24)

This is how the schedule of the pipeline was debugged in round 2: the tile period is the slowest warp's item time plus ~800
cycles, every extra item costs ~850 cycles, thread 0 needs 1800 - 3300 cycles per tile to issue the bulk copies.
"""
import collections
import os
import statistics as st
import sys

KIND = {0: "copy", 1: "scalar", 2: "pack", 3: "zero"}
DT = ["u8", "i8", "u16", "i16", "u32", "i32", "u64", "i64", "f32", "f64"]


def run(path):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pasture_b200 as pb
    from pasture_b200 import algorithms as alg, las
    os.environ["PB200_TILE_TRACE"] = path
    n = 8_000_000
    raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    src = alg.synth_las_fmt0_records(n)
    col = pb.HashMapBuffer(tgt, n, "cuda")
    cv = pb.get_default_las_converter(raw, tgt, scale, offset)

    def section(name, fn):
        fn()  # warm (its launch is logged too; `show` takes the last launch of a section)
        torch.cuda.synchronize()
        open(path, "a").write("== " + name + "\n")
        fn()
        torch.cuda.synchronize()

    open(path, "w").write("")
    section("C2: interleaved raw LAS fmt0 (20 B) -> columnar LasPointFormat0 (35 B)", lambda: cv.convert_into(src, col))
    section("C5 step: C2 + fused AABB", lambda: cv.convert_into_range_with_bounds(src, range(0, n), col, range(0, n)))
    aos = pb.VectorBuffer(tgt, n, "cuda")
    ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
    section("columnar -> interleaved LasPointFormat0 (35 B packed records)", lambda: ident.convert_into(col, aos))
    back = pb.VectorBuffer(raw, n, "cuda")
    wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
    wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, pb.InvScaleOffset(scale, offset), True)
    section("C1 on the GPU, convert semantics (35 B records -> 20 B records)", lambda: wr.convert_into_fresh(aos, back))
    section("LAS egress (columnar 35 B -> fmt0 records + stats)", lambda: las.write_points(col, 0, scale, offset))


def show(path):
    for sec in open(path).read().split("== ")[1:]:
        lines = sec.splitlines()
        starts = [i for i, l in enumerate(lines) if l.startswith("launch")]
        if not starts:
            continue
        L = lines[starts[0]:starts[1]] if len(starts) > 1 else lines[starts[0]:]
        items, T = {}, collections.defaultdict(dict)
        for l in L:
            if l.startswith("warp "):
                w = int(l.split()[1])
                out = []
                for it in l.split("[")[1:]:
                    f = it.rstrip("] ").split()
                    kind, types, xf, nbytes, grp, pts = int(f[1]), f[2], int(f[4]), int(f[6]), int(f[8]), f[10]
                    a, b = types.split("->")
                    what = KIND[kind]
                    if kind == 1:
                        what += f" {DT[int(a)]}->{DT[int(b)]}" + (f" xf{xf}" if xf else "")
                    elif kind in (0, 3):
                        what += f" {nbytes} B" + (f" x{grp}" if grp else "")
                    out.append(f"{what} [{pts}]")
                items[w] = ", ".join(out)
            elif l.startswith("tile "):
                p = l.split()
                T[int(p[1])][int(p[3])] = [int(x) for x in p[4:]]
        nw = len(items)
        period = [min(T[i][w][0] for w in range(nw)) - min(T[i - 1][w][0] for w in range(nw)) for i in range(9, 25)]
        print(f"### {lines[0]}\n\n{L[0]}; tile period (median of tiles 9..24): **{st.median(period):.0f} cycles**\n")
        print("| warp | wait for data | work items | thread 0: issue bulk copies | items (points of the tile) |\n|---|---|---|---|---|")
        for w in range(nw):
            a = [T[i][w][1] - T[i][w][0] for i in range(8, 25)]
            b = [T[i][w][2] - T[i][w][1] for i in range(8, 25)]
            e = [T[i][w][5] - T[i][w][4] if T[i][w][5] else 0 for i in range(8, 25)]
            print(f"| {w} | {st.median(a):.0f} | {st.median(b):.0f} | {st.median(e):.0f} | {items[w]} |")
        print()


if __name__ == "__main__":
    (run if sys.argv[1] == "run" else show)(sys.argv[2])
