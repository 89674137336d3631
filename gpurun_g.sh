#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbvh_query -c 1 -o gpurun_out/prof_knn_r1 python benchmarks/knn_probe.py --points 4000000 --stats 0 > gpurun_out/g_knn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"voxel_reduce|voxel_key|heads_emit" -c 3 -o gpurun_out/prof_voxel_r1 python benchmarks/bench_configs.py --points 50000000 --skip aabb,soa2aos,filter,ransac,las,pnts,c4 > gpurun_out/g_voxel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 4 -c 4 -o gpurun_out/prof_directions_r1 python benchmarks/prof_directions.py --points 20000000 > gpurun_out/g_dir.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/g_knn.log gpurun_out/g_voxel.log gpurun_out/g_dir.log
