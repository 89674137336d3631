#!/bin/bash
mkdir -p gpurun_out
PB200_TUNE_LOG=1 python benchmarks/bench_configs.py --params "convert.autotune=1" --skip aabb,c3,filter,ransac,c4 2> gpurun_out/r3c.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
grep "pb200 tune" gpurun_out/r3c.err | head -150
