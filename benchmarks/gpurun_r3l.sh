#!/bin/bash
mkdir -p gpurun_out
for V in G I J; do
cp benchmarks/build/variants/$V.so pasture_b200/libpasture_b200.so
P=""; [ "$V" = "G" ] && P="convert.cost_item=260"
echo "== variant $V params: $P"
python benchmarks/bench_configs.py --params "$P" --skip aabb,c3,filter,ransac,c4 2> gpurun_out/r3l.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
X=""; [ "$V" = "G" ] && X="--param convert.cost_item=260"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs $X 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds $X 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2 fused', d['ms_per_step'], d['roofline']['frac'])"
done
