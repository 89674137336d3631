#!/bin/bash
# N-GPU evidence: PCIe probe (plain pinned copies, all ranks concurrently) and the bench with e2e.roofline + scaling attribution
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r2n_topo_n$N.txt 2>&1
lscpu | head -30 >> gpurun_out/r2n_topo_n$N.txt 2>&1
free -g >> gpurun_out/r2n_topo_n$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 benchmarks/pcie_probe.py > gpurun_out/r2n_probe_n$N.json 2> gpurun_out/r2n_probe_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
tail -3 gpurun_out/r2n_probe_n$N.err gpurun_out/r2n_bench_n$N.err
python - <<P
import json
p=json.loads(open('gpurun_out/r2n_probe_n$N.json').read().strip().splitlines()[-1])
print('probe', {k:(round(v,2) if isinstance(v,float) else v) for k,v in p.items() if k not in ('topology','ranks')})
print('topo', p['topology'])
d=json.loads(open('gpurun_out/r2n_bench_n$N.json').read().strip().splitlines()[-1])
print('bench', d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['roofline']['frac'])
print('e2e', d['e2e'])
print('attr', d.get('scaling_attribution'))
P
