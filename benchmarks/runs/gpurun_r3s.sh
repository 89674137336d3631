#!/bin/bash
# sharded voxel grid at the GPU counts missing from profiles/sharded_voxel_r2.jsonl (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for A in position las; do
if [ "$N" = "1" ]; then
python benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes $A --check 2>> gpurun_out/sharded_n$N.err | grep '^{' >> gpurun_out/sharded_voxel_n$N.jsonl
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29661 benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes $A --check 2>> gpurun_out/sharded_n$N.err | grep '^{' >> gpurun_out/sharded_voxel_n$N.jsonl
fi
done
tail -3 gpurun_out/sharded_n$N.err; cut -c1-700 gpurun_out/sharded_voxel_n$N.jsonl
