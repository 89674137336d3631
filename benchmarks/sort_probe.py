"""radix sort probe: the two shapes the library uses (C3: 100 M packed voxel keys, 34 bits above 27 index bits, keys only;
C4: 20 M Morton codes, 63 bits, with the point index as payload) -- ms and effective GB/s per pass"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasture_b200.algorithms import radix_sort

def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

g = torch.Generator(device="cuda").manual_seed(1)
n = 100_000_000
keys = (torch.randint(0, 1 << 34, (n,), dtype=torch.int64, device="cuda", generator=g) << 27) | torch.arange(n, dtype=torch.int64, device="cuda")
work = torch.empty_like(keys)
def c3():
    work.copy_(keys); radix_sort(work, None, 27, 61)
copy_ms = timed(lambda: work.copy_(keys))
ms = timed(c3) - copy_ms
print(f"C3 shape: 100M keys-only, 34 bits (5 passes): {ms:.3f} ms  ({(8 + 5 * 16) * n / ms / 1e6:.0f} GB/s algorithmic)", flush=True)
n2 = 20_000_000
codes = torch.randint(0, 1 << 62, (n2,), dtype=torch.int64, device="cuda", generator=g)
idx = torch.arange(n2, dtype=torch.int32, device="cuda")
w2, v2 = torch.empty_like(codes), torch.empty_like(idx)
def c4():
    w2.copy_(codes); v2.copy_(idx); radix_sort(w2, v2, 0, 63)
copy2 = timed(lambda: (w2.copy_(codes), v2.copy_(idx)))
ms2 = timed(c4) - copy2
print(f"C4 shape: 20M (key, payload), 63 bits (8 passes): {ms2:.3f} ms  ({(8 + 8 * 24) * n2 / ms2 / 1e6:.0f} GB/s algorithmic)", flush=True)
ref = torch.sort(keys).values
c3(); print("C3 result equals torch.sort:", bool(torch.equal(work, ref)))
