#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_knn.py -x -q ) 2>&1 | tail -15 > gpurun_out/c_pytest.txt
timeout 600 python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts,c4 > gpurun_out/c_configs.jsonl 2> gpurun_out/c_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches_c3.csv python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts,c4 > /dev/null 2>&1
cat gpurun_out/c_pytest.txt; cat gpurun_out/c_configs.jsonl; tail -3 gpurun_out/c_configs.err
