#!/bin/bash
mkdir -p gpurun_out

python -m pytest tests/test_gpu_voxel.py tests/test_gpu_sort.py tests/test_gpu_knn.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -5 gpurun_out/r2i_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
for o in d['other_configs']:
    print(o['workload'][:30], [round(x,2) for x in o['ms_all']], 'kernel sum', round(o['kernel_ms_sum'],3), 'frac', round(o['frac'],3))
    for k in o['kernels']: print('      ', k['name'], round(k['ms'],3), round(k.get('frac') or 0,3))
P
