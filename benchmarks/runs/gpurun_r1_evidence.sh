#!/bin/bash
# round-1 evidence run: full GPU suite, default bench (both arms), config benchmarks, ncu launch list of the bench
# command, one full capture of the top kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r1.txt
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1.err
python benchmarks/bench_configs.py > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 3 -c 1 -o gpurun_out/prof_convert_r1f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/smoke_r1.txt gpurun_out/pytest_gpu_r1.txt; cat gpurun_out/bench_r1.json; cat gpurun_out/bench_r1_reference.json; cat gpurun_out/configs_r1.jsonl
