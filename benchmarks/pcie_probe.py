#!/usr/bin/env python
"""pcie_probe.py -- the ceiling of the host-buffer (e2e) path: plain pinned cudaMemcpyAsync, no library code.

What `bench.py`'s e2e leg moves per step and rank is 2.0 GB host->device (raw LAS records) and 3.5 GB device->host
(the ten columns).  This probe copies exactly those byte counts between pinned host memory and HBM with
torch's `copy_(non_blocking=True)` (= cudaMemcpyAsync on a copy engine), in the same process layout as the bench
(one process per GPU under torchrun, all ranks at the same time), and reports

  h2d_alone / d2h_alone   one direction at a time, all ranks concurrently
  both                    H2D on one stream and D2H on another at the same time, all ranks concurrently; its duration is
                          the FLOOR of an e2e step (the converter can at best hide its kernel behind these copies)
  solo                    the same on rank 0 while the other ranks idle (what one PCIe link can do without host contention)

so that an e2e number can be read against the hardware: e2e.roofline.frac = floor_ms / e2e_ms.

  python benchmarks/pcie_probe.py
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/pcie_probe.py [--no-bind]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_topology(dev_index):
    """NUMA node / local CPUs of the GPU (sysfs) and what this process may run on"""
    import torch
    info = {"allowed_cpus": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
    try:
        p = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["pci"] = bus
        base = "/sys/bus/pci/devices/" + bus
        for key in ("numa_node", "local_cpulist"):
            try:
                with open(os.path.join(base, key)) as fh:
                    info[key] = fh.read().strip()
            except OSError:
                info[key] = None
    except Exception as exc:  # noqa: BLE001
        info["error"] = str(exc)
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        info["numa_nodes"] = len(nodes)
    except OSError:
        info["numa_nodes"] = None
    return info


def _timed(fn, reps, sync, barrier):
    """best wall-clock of `reps` runs, every run bracketed by a barrier (all ranks start together) and a device sync"""
    best = None
    for _ in range(reps):
        sync()
        barrier()
        t0 = time.perf_counter()
        fn()
        sync()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best


def probe(dev, h2d_bytes, d2h_bytes, reps=3, dist=None, solo=True, chunk_bytes=0):
    """returns {'h2d_alone_gbs','d2h_alone_gbs','both_ms','both_h2d_gbs','both_d2h_gbs', 'solo_*'} with per-rank times
    reduced by MAX over ranks (the slowest link sets an e2e step).  chunk_bytes > 0 splits every copy into chunks of
    that size (what a staged pipeline issues)."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)  # touch
    h_out.fill_(0)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(d2h_bytes, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy(dst, src, stream):
        with torch.cuda.stream(stream):
            if chunk_bytes and chunk_bytes < src.numel():
                for o in range(0, src.numel(), chunk_bytes):
                    dst[o:o + chunk_bytes].copy_(src[o:o + chunk_bytes], non_blocking=True)
            else:
                dst.copy_(src, non_blocking=True)

    def sync():
        torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"world": world, "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes, "chunk_bytes": chunk_bytes}
    copy(d_in, h_in, s_in); copy(h_out, d_out, s_out); sync()  # warm-up
    t = maxr(_timed(lambda: copy(d_in, h_in, s_in), reps, sync, barrier))
    out["h2d_alone_gbs"] = h2d_bytes / t / 1e9
    t = maxr(_timed(lambda: copy(h_out, d_out, s_out), reps, sync, barrier))
    out["d2h_alone_gbs"] = d2h_bytes / t / 1e9
    t = maxr(_timed(lambda: (copy(d_in, h_in, s_in), copy(h_out, d_out, s_out)), reps, sync, barrier))
    out["both_ms"] = t * 1e3
    out["both_h2d_gbs"] = h2d_bytes / t / 1e9
    out["both_d2h_gbs"] = d2h_bytes / t / 1e9
    if solo and world > 1:  # rank 0 alone: the other ranks wait at the barrier behind it
        if rank == 0:
            t0 = _timed(lambda: (copy(d_in, h_in, s_in), copy(h_out, d_out, s_out)), reps, sync, lambda: None)
        else:
            t0 = 0.0
        dist.barrier()
        out["solo_both_ms"] = maxr(t0) * 1e3
    del h_in, h_out, d_in, d_out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-bytes", type=int, default=2_000_000_000)
    ap.add_argument("--d2h-bytes", type=int, default=3_500_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunk-mb", type=int, default=0)
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process to the GPU's NUMA-local CPUs first")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        d = dist
    topo = host_topology(dev.index)
    bound = None
    if not args.no_bind:
        import pasture_b200 as pb
        bound = pb.get_context(dev.index).bind_host_thread()
    res = probe(dev, args.h2d_bytes, args.d2h_bytes, args.reps, d, chunk_bytes=args.chunk_mb << 20)
    res["bound"] = bound
    res["topology"] = topo
    if world > 1:
        topos = [None] * world
        dist.all_gather_object(topos, {"topo": topo, "bound": bound})
        res["ranks"] = topos
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
