#!/bin/bash
# N-GPU evidence of round 2: bench (probe + e2e.roofline + scaling attribution), sharded voxel grid (positions / LAS attributes,
# checked against the single-device filter over the whole cloud), C4 replicas-only normals
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:3}" > gpurun_out/$2_n$N.json 2> gpurun_out/$2_n$N.err; grep -v "OMP_NUM\|\*\*\*\|^$\|NCCL version" gpurun_out/$2_n$N.err | tail -n 2; tail -c 2500 gpurun_out/$2_n$N.json; echo; }
run 29641 r2r_bench bench.py --gpus $N --steps 50 --warmup 3
run 29642 r2r_sharded_position benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes position --check
run 29643 r2r_sharded_las benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes las --check
run 29644 r2r_normals benchmarks/sharded_normals.py --points 100000000
