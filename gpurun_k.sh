#!/bin/bash
mkdir -p gpurun_out
for mb in 32 64 128 256 512; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --param convert.stage_chunk_mb=$mb > gpurun_out/k_$mb.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/k_$mb.json').read().strip().splitlines()[-1]); print($mb, round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/1e9,3))"
done
python - <<PY
import torch,time
a=torch.empty(1<<30,dtype=torch.uint8,device="cuda"); h=torch.empty(1<<30,dtype=torch.uint8).pin_memory()
for name,fn in (("d2h",lambda: h.copy_(a,non_blocking=True)),("h2d",lambda: a.copy_(h,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, round(5*(1<<30)/(time.perf_counter()-t)/1e9,1),"GB/s")
PY
