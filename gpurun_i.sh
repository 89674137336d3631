#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/i_bench_n2_peer.json 2> gpurun_out/i_bench_n2_peer.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --nccl-bounds > gpurun_out/i_bench_n2_nccl.json 2> gpurun_out/i_bench_n2_nccl.err
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -5 > gpurun_out/i_pytest.txt
for f in gpurun_out/i_bench_n2_peer.json gpurun_out/i_bench_n2_nccl.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['config']['bounds_exchange'][:40], d['gpu_launches'])"; done
cat gpurun_out/i_pytest.txt; tail -n 3 gpurun_out/i_bench_n2_peer.err
