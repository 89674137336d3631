#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 > gpurun_out/d_pytest.txt
timeout 600 python benchmarks/bench_configs.py --skip aabb,soa2aos,las,pnts,c4 > gpurun_out/d_configs.jsonl 2> gpurun_out/d_configs.err
cat gpurun_out/d_pytest.txt; cat gpurun_out/d_configs.jsonl; tail -3 gpurun_out/d_configs.err
