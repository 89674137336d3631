#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_convert.py tests/test_gpu_las_io.py tests/test_gpu_pnts.py -x -q 2>&1 | tail -5
for V in A D E; do
cp benchmarks/build/variants/$V.so pasture_b200/libpasture_b200.so
echo "== variant $V"
python benchmarks/bench_configs.py --skip aabb,c3,filter,ransac,c4 2> gpurun_out/r3f.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs 2>> gpurun_out/r3f.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2', d['ms_per_step'], d['roofline']['frac'])"
done
tail -3 gpurun_out/r3f.err
