#!/bin/bash
mkdir -p gpurun_out
PB200_TUNE_LOG=1 python benchmarks/bench_configs.py --skip aabb,c3,filter,ransac,c4,las,pnts 2> gpurun_out/r3t.err | tee gpurun_out/r3t.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:70])"
grep "default\|\*" gpurun_out/r3t.err | head -20
