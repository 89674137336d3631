#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_knn.py -x -q ) 2>&1 | tail -4 > gpurun_out/j_pytest.txt
timeout 600 python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts > gpurun_out/j_configs.jsonl 2> gpurun_out/j_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j_launches_c3.csv python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts,c4 > /dev/null 2>&1
cat gpurun_out/j_pytest.txt; cat gpurun_out/j_configs.jsonl
