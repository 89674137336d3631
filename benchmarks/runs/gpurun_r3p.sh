#!/bin/bash
# round-2 evidence run #2 (one GPU), after the converter's schedule / issue-order / packed-record work:
# smoke, full GPU suite, both bench arms, config rows, launch list, per-kernel ncu table, C2 / C5 source-level captures
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.txt 2>&1; tail -2 gpurun_out/smoke_r2.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r2.txt; cat gpurun_out/pytest_gpu_r2.txt
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err; cut -c1-900 gpurun_out/bench_r2_n1.json
python bench.py --impl reference > gpurun_out/bench_r2_n1_reference.json 2>> gpurun_out/bench_r2_n1.err; cut -c1-300 gpurun_out/bench_r2_n1_reference.json
python benchmarks/bench_configs.py > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err; tail -3 gpurun_out/configs_r2.err
python - <<'PY'
import json
for l in open("gpurun_out/configs_r2.jsonl"):
    d = json.loads(l)
    print(round(d.get("ms", 0), 3), round(d.get("frac_of_measured_peak") or 0, 3), d.get("frac_of_measured_fp64_rate"), (d.get("clocks") or {}).get("sm_mhz"), d["config"][:80])
PY
K='^(convert_|bounds_|minmax_|morton_|voxel_|radix_|heads_|lbvh_|gather_pos|reproject_|filter_c|ransac_r|qpos)'
ncu --set full --clock-control none -k regex:"$K" -c 110 -o /tmp/prof_all_r2 python benchmarks/prof_all.py --points 50000000 > gpurun_out/r3p_prof.log 2>&1
tail -2 gpurun_out/r3p_prof.log
python benchmarks/ncu_summary.py /tmp/prof_all_r2.ncu-rep --all --bytes-json gpurun_out/prof_all_bytes.json --out gpurun_out/ncu_all_r2.json \
  --command "ncu --set full --clock-control none -k regex:$K -c 110 python benchmarks/prof_all.py --points 50000000" \
  --workload "every kernel family once, 50 M points (kNN: 4 M)" > gpurun_out/r3p_summary.txt 2>&1
tail -5 gpurun_out/r3p_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r3p_bench_under_ncu.json 2>&1
rm -f gpurun_out/prof_convert_c2_r2.ncu-rep gpurun_out/prof_convert_c5_r2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_c2_r2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_c5_r2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
