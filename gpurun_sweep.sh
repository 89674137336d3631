python -m pytest tests/test_gpu_convert.py -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for cfg in "" "--param convert.threads=512 --param convert.ctas_per_sm=1" "--param convert.threads=512 --param convert.ctas_per_sm=1 --param convert.stages=3" "--param convert.threads=384 --param convert.ctas_per_sm=1"  "--param convert.threads=512 --param convert.ctas_per_sm=2" "--param convert.threads=384 --param convert.ctas_per_sm=2" "--param convert.threads=256 --param convert.ctas_per_sm=3" "--param convert.threads=512 --param convert.ctas_per_sm=1 --fused-bounds"; do
  echo "== $cfg"; $B $cfg 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['roofline']['best_launch_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
done
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_r1d python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --param convert.threads=512 --param convert.ctas_per_sm=1 > /dev/null 2>&1
