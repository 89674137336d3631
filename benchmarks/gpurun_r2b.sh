#!/bin/bash
# round 2, voxel pipeline rework: parity of the fused emit+reduce kernel, 9-bit digit passes and the locality pass; C3 timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_voxel.py tests/test_gpu_sort.py tests/test_gpu_knn.py tests/test_gpu_multigpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.txt
cat gpurun_out/r2b_pytest.txt
python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python bench.py --steps 20 --no-e2e --no-cpu-baseline --param voxel.partition=0 > gpurun_out/r2b_bench_nopart.json 2>> gpurun_out/r2b_bench.err
python bench.py --steps 20 --no-e2e --no-cpu-baseline --param voxel.partition=0 --param sort.force_8bit=1 > gpurun_out/r2b_bench_nopart_8bit.json 2>> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_bench.err
python - <<'P'
import json
for f in ('r2b_bench','r2b_bench_nopart','r2b_bench_nopart_8bit'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(f, d['ms_per_step'], d['roofline']['frac'])
    for o in d.get('other_configs') or []:
        print('  ', o['workload'][:40], round(o['ms'],3), 'kernel sum', round(o['kernel_ms_sum'],3), o['clocks']['sm_mhz'])
        for k in o['kernels']: print('      ', k['name'], round(k['ms'],3), k.get('launch_groups'), round(k.get('frac') or 0,3))
P
