#!/bin/bash
mkdir -p gpurun_out
run() { python benchmarks/bench_configs.py --params "$1" --skip "$2" 2>> gpurun_out/r2z.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"; }
for d in 4 8 12 16 24; do echo "cost_div=$d"; run "convert.cost_div=$d" aabb,c3,filter,ransac,c4,pnts; done
for b in 4 16 32; do for s in 1 3 5; do echo "group base=$b store=$s"; run "convert.cost_group_base=$b,convert.cost_group_store=$s" aabb,c3,filter,ransac,c4,pnts,las; done; done
for pb_ in 12 24 40; do for ps in 6 12 18; do echo "pack base=$pb_ per_src=$ps"; run "convert.cost_pack_base=$pb_,convert.cost_pack_per_src=$ps" aabb,c3,filter,ransac,c4,pnts,soa2aos; done; done
tail -3 gpurun_out/r2z.err
