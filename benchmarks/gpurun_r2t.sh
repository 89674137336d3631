#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_knn.py -x -q 2>&1 | tail -3
python - <<'P'
import time, torch, pasture_b200 as pb
from pasture_b200 import algorithms as alg
ctx = pb.get_context()
for n in (20_000_000, 100_000_000):
    src = alg.synth_terrain_positions(n)
    alg.compute_normals(src, 16); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter(); alg.compute_normals(src, 16); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
    ctx.profile(True); alg.compute_normals(src, 16); ph = ctx.profile_read(); ctx.profile(False)
    print(n, round(best, 2), 'ms', [(k, round(v, 2)) for k, v in ph if k.startswith('knn.q')])
    del src
P
