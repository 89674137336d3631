#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_knn.py tests/test_gpu_convert.py tests/test_gpu_las_io.py -x -q 2>&1 | tail -8
python - <<'P'
import time, torch, pasture_b200 as pb
from pasture_b200 import algorithms as alg
n = 20_000_000
src = alg.synth_terrain_positions(n)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
    return best
print('normals 20M full', round(t(lambda: alg.compute_normals(src, 16)), 2), 'ms')
for g in (2, 4, 8):
    print(f'normals 20M, one range of 1/{g}', round(t(lambda: alg.compute_normals(src, 16, query_range=range(n // g, 2 * (n // g)))), 2), 'ms')
P
