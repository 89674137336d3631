#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_convert.py tests/test_gpu_las_io.py tests/test_gpu_pnts.py tests/test_gpu_filter.py -x -q ) 2>&1 | tail -8 > gpurun_out/h_pytest.txt
timeout 600 python benchmarks/bench_configs.py --skip aabb,c3,filter,ransac,pnts,c4 > gpurun_out/h_configs.jsonl 2> gpurun_out/h_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 1 -c 1 -o gpurun_out/prof_soa2aos_r1 python benchmarks/prof_directions.py --points 20000000 > gpurun_out/h_dir.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/h_launches_las.csv python benchmarks/bench_configs.py --points 20000000 --skip aabb,c3,filter,ransac,pnts,c4,soa2aos > /dev/null 2>&1
cat gpurun_out/h_pytest.txt; cat gpurun_out/h_configs.jsonl
