#!/bin/bash
mkdir -p gpurun_out
for I in 120 160 200 260 320; do
echo "== cost_item=$I"
python benchmarks/bench_configs.py --params "convert.cost_item=$I" --skip aabb,c3,filter,ransac,c4 2>> gpurun_out/r3k.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs --param convert.cost_item=$I 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds --param convert.cost_item=$I 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2 fused', d['ms_per_step'], d['roofline']['frac'])"
done
