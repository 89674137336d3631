#!/bin/bash
# per-kernel ncu evidence (round 2): ONE --set full run over every kernel family, summarised on the box (the raw report is
# too large to travel: 64 MiB limit), plus small source-level reports of the four kernels that dominate C2/C3/C4
mkdir -p gpurun_out
K='^(convert_|bounds_|minmax_|morton_|voxel_|radix_|heads_|lbvh_|gather_pos|reproject_|filter_c|ransac_r|qpos)'
ncu --set full --clock-control none -k regex:"$K" -c 110 -o /tmp/prof_all_r2 python benchmarks/prof_all.py --points 50000000 > gpurun_out/r2o_prof.log 2>&1
tail -2 gpurun_out/r2o_prof.log
python benchmarks/ncu_summary.py /tmp/prof_all_r2.ncu-rep --all --bytes-json gpurun_out/prof_all_bytes.json --out gpurun_out/ncu_all_r2.json \
  --command "ncu --set full --clock-control none -k regex:$K -c 110 python benchmarks/prof_all.py --points 50000000" \
  --workload "every kernel family once, 50 M points (kNN: 4 M)" > gpurun_out/r2o_summary.txt 2>&1
cat gpurun_out/r2o_summary.txt
ls -la /tmp/prof_all_r2.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_under_ncu.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_c2_r2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_c5_r2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --fused-bounds > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
