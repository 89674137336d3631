"""sweep of the convert kernel's cost-model constants on the two directions that use them (LAS egress: pack + division,
C1 on the GPU: division)"""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb
from pasture_b200 import algorithms as alg, las
from pasture_b200.context import get_context
ctx = get_context()
n = 50_000_000
raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
src = alg.synth_las_fmt0_records(n)
col = pb.HashMapBuffer(tgt, n, "cuda")
cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
cv.convert_into(src, col)
aos = pb.VectorBuffer(tgt, n, "cuda")
pb.BufferLayoutConverter.for_layouts(tgt, tgt).convert_into(col, aos)
back = pb.VectorBuffer(raw, n, "cuda")
def timed(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for div, pbase, psrc in [(24, 6, 4), (24, 12, 8), (24, 24, 12), (24, 36, 16), (24, 48, 24), (24, 64, 32), (32, 36, 16), (16, 24, 12)]:
    ctx.set_param("convert.cost_div", div); ctx.set_param("convert.cost_pack_base", pbase); ctx.set_param("convert.cost_pack_per_src", psrc)
    wr = pb.BufferLayoutConverter.for_layouts_with_default(tgt, raw)
    wr.set_custom_mapping_with_transformation(pb.attributes.POSITION_3D, pb.ATTRIBUTE_LOCAL_LAS_POSITION, pb.InvScaleOffset(0.001, (500000.0, 5400000.0, 100.0)), True)
    c1 = timed(lambda: wr.convert_into(aos, back))
    eg = timed(lambda: las.write_points(col, 0, (0.001,) * 3, (500000.0, 5400000.0, 100.0)))
    print(f"div={div} pack={pbase}+{psrc}/src: C1-on-GPU {c1:.3f} ms, LAS egress {eg:.3f} ms (50M points)", flush=True)
