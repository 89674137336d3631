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
ident = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
dst2 = pb.HashMapBuffer(tgt, n, "cuda")
for cbase, cstore in [(6, 1), (4, 1), (8, 1), (12, 1), (6, 2), (6, 3), (10, 2), (3, 1)]:
    ctx.set_param("convert.cost_copy_base", cbase); ctx.set_param("convert.cost_store", cstore)
    cv2 = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
    ident2 = pb.BufferLayoutConverter.for_layouts(tgt, tgt)
    c2 = timed(lambda: cv2.convert_into(src, dst2))
    s2a = timed(lambda: ident2.convert_into(col, aos))
    eg = timed(lambda: las.write_points(col, 0, (0.001,) * 3, (500000.0, 5400000.0, 100.0)))
    print(f"copy_base={cbase} store={cstore}: C2 {c2:.4f} ms, columnar->interleaved {s2a:.3f} ms, LAS egress {eg:.3f} ms (50M points)", flush=True)
