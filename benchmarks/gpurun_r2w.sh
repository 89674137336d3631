#!/bin/bash
# tile-kernel shape sweep for the non-headline conversion directions
mkdir -p gpurun_out
: > gpurun_out/r2w_sweep.jsonl
for P in "convert.threads=512" "convert.threads=768" "convert.threads=1024" "convert.threads=512,convert.ctas_per_sm=2" "convert.threads=1024,convert.stages=1"; do
  echo "== $P"
  python benchmarks/bench_configs.py --params "$P" --skip aabb,c3,filter,ransac,c4 2>> gpurun_out/r2w.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['params'] = '$P'
    open('gpurun_out/r2w_sweep.jsonl','a').write(json.dumps(d)+'\n')
    print(round(d['ms'],3), round(d['frac_of_measured_peak'] or 0,3), d['config'][:70])"
  python bench.py --steps 30 --no-e2e --no-cpu-baseline --no-other-configs $(for kv in ${P//,/ }; do echo -n "--param $kv "; done) 2>> gpurun_out/r2w.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2', d['ms_per_step'], d['roofline']['frac'])"
done
tail -5 gpurun_out/r2w.err
