#!/bin/bash
# 2-GPU evidence: bench at N=2 with the peer-memory exchange and with NCCL, sharded voxel grid with the single-device
# check, the multi-GPU tests
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 100 --warmup 3 --no-e2e > gpurun_out/i_bench_n2_peer.json 2> gpurun_out/i_bench_n2_peer.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 100 --warmup 3 --no-e2e --nccl-bounds > gpurun_out/i_bench_n2_nccl.json 2> gpurun_out/i_bench_n2_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 benchmarks/sharded_voxel.py --points-per-gpu 50000000 --check > gpurun_out/i_sharded_n2.json 2> gpurun_out/i_sharded_n2.err
timeout 600 python benchmarks/sharded_voxel.py --points-per-gpu 100000000 > gpurun_out/i_sharded_n1.json 2> gpurun_out/i_sharded_n1.err
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3 > gpurun_out/i_pytest.txt
for f in gpurun_out/i_bench_n2_peer.json gpurun_out/i_bench_n2_nccl.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['config']['bounds_exchange'][:30], d['gpu_launches'], d['clocks'])"; done
cat gpurun_out/i_sharded_n2.json gpurun_out/i_sharded_n1.json gpurun_out/i_pytest.txt
