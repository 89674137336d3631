#!/bin/bash
# multi-GPU bench line (both arms) with the final converter; N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
tail -2 gpurun_out/bench_r2_n$N.err; cut -c1-700 gpurun_out/bench_r2_n$N.json
python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -2
