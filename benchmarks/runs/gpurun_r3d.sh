#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; python benchmarks/bench_configs.py --params "$1" --skip aabb,c3,filter,ransac,c4,pnts 2>> gpurun_out/r3d.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"; }
run "convert.tile_points=2048"
run "convert.tile_points=1024"
run "convert.tile_points=512"
run "convert.tile_points=1024,convert.stages=4"
run "convert.tile_points=1024,convert.threads=256"
run "convert.tile_points=1024,convert.ctas_per_sm=2"
run "convert.tile_points=1024,convert.ctas_per_sm=2,convert.threads=256"
run "convert.tile_points=512,convert.ctas_per_sm=4,convert.threads=256"
run "convert.tile_points=512,convert.ctas_per_sm=4,convert.threads=128"
run "convert.stages=3"
tail -3 gpurun_out/r3d.err
