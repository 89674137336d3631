#!/bin/bash
mkdir -p gpurun_out
python benchmarks/bench_configs.py --skip aabb,c3,soa2aos,filter,ransac,las,pnts > gpurun_out/configs_c4.jsonl 2> gpurun_out/configs_c4.err; cut -c1-300 gpurun_out/configs_c4.jsonl
cp benchmarks/build/variants/T.so pasture_b200/libpasture_b200.so
python benchmarks/tile_trace.py run gpurun_out/tile_trace_r2.txt; wc -l gpurun_out/tile_trace_r2.txt
