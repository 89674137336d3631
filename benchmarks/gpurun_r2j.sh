#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_convert.py -x -q -k fresh 2>&1 | tail -8
python benchmarks/bench_configs.py --skip aabb,c3,filter,ransac,pnts,c4,las 2> gpurun_out/r2k_configs.err | tee gpurun_out/r2k_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:110])"
tail -3 gpurun_out/r2k_configs.err
