#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m pytest tests/test_gpu_voxel.py tests/test_gpu_reduce.py tests/test_gpu_knn.py tests/test_gpu_convert.py -x -q 2>&1 | tail -8
for attrs in position las; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 benchmarks/sharded_voxel.py --points-per-gpu 20000000 --attributes $attrs --check > gpurun_out/r2p_sharded_${attrs}_n$N.json 2> gpurun_out/r2p_sharded_${attrs}_n$N.err
  tail -n 3 gpurun_out/r2p_sharded_${attrs}_n$N.err; cat gpurun_out/r2p_sharded_${attrs}_n$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 benchmarks/sharded_voxel.py --points-per-gpu 100000000 --attributes position > gpurun_out/r2p_sharded_100M_n$N.json 2> gpurun_out/r2p_sharded_100M_n$N.err
tail -n 3 gpurun_out/r2p_sharded_100M_n$N.err; cat gpurun_out/r2p_sharded_100M_n$N.json
