"""helper for ncu: runs the non-headline conversion directions once each at --points"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb
from pasture_b200 import algorithms as alg
ap = argparse.ArgumentParser(); ap.add_argument("--points", type=int, default=20_000_000); args = ap.parse_args()
n = args.points
raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
src = alg.synth_las_fmt0_records(n)
col = pb.HashMapBuffer(tgt, n, "cuda")
cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
cv.convert_into(src, col)
aos = pb.VectorBuffer(tgt, n, "cuda")
ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
for _ in range(3):
    ident.convert_into(col, aos)      # columnar -> interleaved (packed 35 B)
back = pb.VectorBuffer(raw, n, "cuda")
wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION,
                                          pb.InvScaleOffset(0.001, (500000.0, 5400000.0, 100.0)), True)
for _ in range(3):
    wr.convert_into(aos, back)        # interleaved 35 B -> interleaved 20 B (write direction)
torch.cuda.synchronize()
