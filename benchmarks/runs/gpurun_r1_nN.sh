#!/bin/bash
# N-GPU evidence (N = number of visible GPUs): bench with the peer-memory exchange
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/m_bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['config']['bounds_exchange'][:30], d['gpu_launches'], d['clocks'])"
tail -n 2 gpurun_out/m_bench_n$N.err
