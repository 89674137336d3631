#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_convert.py tests/test_gpu_las_io.py tests/test_gpu_pnts.py tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
for P in "" "convert.cost_warp0=0" "convert.cost_warp0=120" "convert.cost_warp0=400" "convert.autotune=1"; do
echo "== params: $P"
PB200_TUNE_LOG=1 python benchmarks/bench_configs.py --params "$P" --skip aabb,c3,filter,ransac,c4 2> gpurun_out/r3n.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:60])"
grep "default\|\*" gpurun_out/r3n.err | head -30
X=""; for kv in ${P//,/ }; do X="$X --param $kv"; done
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs $X 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds $X 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   C2 fused', d['ms_per_step'], d['roofline']['frac'])"
done
