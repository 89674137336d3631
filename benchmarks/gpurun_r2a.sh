#!/bin/bash
# round-2 first evidence run (1 GPU): tests, bench, gather probe, ncu of every kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2a_pytest.txt
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
benchmarks/build/gather_probe > gpurun_out/r2a_gather.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(convert|bounds|minmax|morton|voxel|radix|heads|lbvh|gather_pos|reproject|filter_c|ransac_r|return_hist|pnts)' -c 90 -o gpurun_out/prof_all_r2a python benchmarks/prof_all.py --points 50000000 > gpurun_out/r2a_prof.log 2>&1
tail -3 gpurun_out/r2a_prof.log
cat gpurun_out/r2a_pytest.txt; tail -c 1500 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_gather.jsonl | head -40
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e'])
for o in d.get('other_configs') or []:
    print(o['workload'][:40], o['ms'], o['ms_all'], o['kernel_ms_sum'], o['clocks'])
    for k in o['kernels']: print('   ', k)
P
