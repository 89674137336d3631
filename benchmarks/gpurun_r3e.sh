#!/bin/bash
mkdir -p gpurun_out
cp benchmarks/build/variants/T.so pasture_b200/libpasture_b200.so
rm -f gpurun_out/tile_trace.txt
PB200_TILE_TRACE=gpurun_out/tile_trace.txt python - <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
import pasture_b200 as pb
from pasture_b200 import algorithms as alg
n = 8_000_000
raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
src = alg.synth_las_fmt0_records(n)
col = pb.HashMapBuffer(tgt, n, "cuda")
cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
open("gpurun_out/tile_trace.txt", "a").write("== C2\n")
cv.convert_into(src, col); torch.cuda.synchronize()
aos = pb.VectorBuffer(tgt, n, "cuda")
ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
open("gpurun_out/tile_trace.txt", "a").write("== soa2aos\n")
ident.convert_into(col, aos); torch.cuda.synchronize()
back = pb.VectorBuffer(raw, n, "cuda")
wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, pb.InvScaleOffset(0.001, (500000.0, 5400000.0, 100.0)), True)
open("gpurun_out/tile_trace.txt", "a").write("== c1 fresh\n")
wr.convert_into_fresh(aos, back); torch.cuda.synchronize()
from pasture_b200 import las
open("gpurun_out/tile_trace.txt", "a").write("== egress\n")
las.write_points(col, 0, (0.001,) * 3, (500000.0, 5400000.0, 100.0)); torch.cuda.synchronize()
PY
python - <<PY2
PY2
wc -l gpurun_out/tile_trace.txt
