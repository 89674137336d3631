#!/bin/bash
mkdir -p gpurun_out
for P in "convert.cost_track=0" "convert.cost_track=2" "convert.cost_track=4" "convert.cost_track=8" "convert.cost_track=4,convert.cost_warp0=120"; do
echo "== params: $P"
python benchmarks/bench_configs.py --params "$P" --skip aabb,c3,filter,ransac,c4,soa2aos,pnts 2> gpurun_out/r3o.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
X=""; for kv in ${P//,/ }; do X="$X --param $kv"; done
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds $X 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2 fused', d['ms_per_step'], d['roofline']['frac'])"
done
