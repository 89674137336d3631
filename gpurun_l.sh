#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_convert.py tests/test_gpu_las_io.py tests/test_gpu_pnts.py tests/test_gpu_multigpu.py -x -q ) 2>&1 | tail -4
python bench.py --no-e2e --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2', d['ms_per_step'], d['roofline']['frac'])"
python benchmarks/bench_configs.py --skip aabb,c3,filter,ransac,c4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms'],3), round(d['frac_of_measured_peak'],3), d['config'][:70])"
