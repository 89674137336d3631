#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e_pytest.txt
cat gpurun_out/r2e_pytest.txt
python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -5 gpurun_out/r2e_bench.err
python - <<'P'
import json
for f in ('r2e_bench',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(d['ms_per_step'], d['roofline']['frac'])
    for o in d['other_configs']:
        print(f, o['workload'][:30], [round(x,2) for x in o['ms_all']], 'kernel sum', round(o['kernel_ms_sum'],3), 'call profiled', o['call_ms_profiled'], 'frac', round(o['frac'],3))
        for k in o['kernels']: print('      ', k['name'], round(k['ms'],3), round(k.get('frac') or 0,3))
        for t in o['host_timeline_us']: print('   host', t)
        print('   pool', o.get('pool_MB'))
P
