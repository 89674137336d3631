#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --no-e2e --no-cpu-baseline --param voxel.partition=0 > gpurun_out/r2d_bench_nopart.json 2> gpurun_out/r2d_bench.err
python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench.json 2>> gpurun_out/r2d_bench.err
tail -5 gpurun_out/r2d_bench.err
python - <<'P'
import json
for f in ('r2d_bench_nopart','r2d_bench'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    o=d['other_configs'][0]
    print(f, [round(x,2) for x in o['ms_all']], 'kernel sum', round(o['kernel_ms_sum'],3), 'call profiled', o['call_ms_profiled'])
    for k in o['kernels']: print('      ', k['name'], round(k['ms'],3))
    for t in o['host_timeline_us']: print('   host', t)
    print('   pool', o.get('pool_MB'))
P
