#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_voxel.py -x -q ) 2>&1 | tail -15 > gpurun_out/e_pytest.txt
timeout 600 python benchmarks/sharded_voxel.py --points-per-gpu 20000000 --check > gpurun_out/e_sharded1.json 2> gpurun_out/e_sharded1.err
cat gpurun_out/e_pytest.txt; cat gpurun_out/e_sharded1.json; tail -5 gpurun_out/e_sharded1.err
