#!/bin/bash
# 2-GPU evidence: NCCL bench at N=2, sharded voxel grid with the single-device check, the 2-rank test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 benchmarks/sharded_voxel.py --points-per-gpu 50000000 --check > gpurun_out/f_sharded_n2.json 2> gpurun_out/f_sharded_n2.err
timeout 600 python benchmarks/sharded_voxel.py --points-per-gpu 100000000 > gpurun_out/f_sharded_n1.json 2> gpurun_out/f_sharded_n1.err
( timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q ) 2>&1 | tail -5 > gpurun_out/f_pytest.txt
cat gpurun_out/f_gpus.txt gpurun_out/f_bench_n2.json gpurun_out/f_sharded_n2.json gpurun_out/f_sharded_n1.json gpurun_out/f_pytest.txt; tail -3 gpurun_out/f_bench_n2.err gpurun_out/f_sharded_n2.err
