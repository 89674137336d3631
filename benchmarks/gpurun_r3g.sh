#!/bin/bash
mkdir -p gpurun_out
for V in TA TD; do
cp benchmarks/build/variants/$V.so pasture_b200/libpasture_b200.so
rm -f gpurun_out/tile_trace_$V.txt
PB200_TILE_TRACE=gpurun_out/tile_trace_$V.txt python - <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
import pasture_b200 as pb
from pasture_b200 import algorithms as alg
f = os.environ["PB200_TILE_TRACE"]
n = 8_000_000
raw, tgt = pb.PointLayout.las_raw(0), pb.PointLayout.las_default(0)
src = alg.synth_las_fmt0_records(n)
col = pb.HashMapBuffer(tgt, n, "cuda")
cv = pb.get_default_las_converter(raw, tgt, (0.001,) * 3, (500000.0, 5400000.0, 100.0))
cv.convert_into(src, col); torch.cuda.synchronize()
open(f, "a").write("== C2\n")
cv.convert_into(src, col); torch.cuda.synchronize()
from pasture_b200 import tiles3d
_l = pb.PointLayout.from_attributes([pb.attributes.POSITION_3D, pb.attributes.COLOR_RGB])
s2 = pb.HashMapBuffer(_l, n, "cuda")
s2.columns[0][: 24 * n].view(torch.float64).uniform_(-1000.0, 1000.0)
w = tiles3d.PntsWriter(_l)
w.write(s2); torch.cuda.synchronize()
open(f, "a").write("== pnts\n")
w.write(s2); torch.cuda.synchronize()
PY
done
wc -l gpurun_out/tile_trace_T*.txt
