#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lbvh_query -c 1 -o gpurun_out/prof_knn_query_r2 python benchmarks/prof_all.py --points 50000000 --knn-points 4000000 --only knn > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 2 -c 4 -o gpurun_out/prof_directions_r2 python benchmarks/prof_all.py --points 50000000 --only convert > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_emit_reduce -c 1 -o gpurun_out/prof_voxel_emit_r2 python benchmarks/prof_all.py --points 50000000 --only reduce > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
