#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for attrs in position las; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes $attrs $CHECK > gpurun_out/r2q_sharded_100M_${attrs}_n$N.json 2> gpurun_out/r2q_sharded_100M_${attrs}_n$N.err
grep -v "OMP_NUM\|\*\*\*\|^$" gpurun_out/r2q_sharded_100M_${attrs}_n$N.err | tail -n 3; cat gpurun_out/r2q_sharded_100M_${attrs}_n$N.json
done
