#!/bin/bash
# session-2 evidence run A: GPU suite, config benchmarks, per-kernel launch lists of C3 / C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 > gpurun_out/a_pytest.txt
python benchmarks/bench_configs.py > gpurun_out/a_configs.jsonl 2> gpurun_out/a_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches_c3.csv python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts,c4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches_c4.csv python benchmarks/bench_configs.py --skip aabb,soa2aos,filter,ransac,las,pnts,c3 > /dev/null 2>&1
cat gpurun_out/a_pytest.txt; cat gpurun_out/a_configs.jsonl; tail -3 gpurun_out/a_configs.err
