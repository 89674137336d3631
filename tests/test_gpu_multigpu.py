"""Multi-GPU path on the GPU box: S logical shards on ONE GPU + the host-side merge must equal the single-shot
result (the analogue of testing a cluster without a cluster), and -- when >= 2 GPUs are visible --
a real 2-rank NCCL run of bench.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle as O
import pasture_b200 as pb
from pasture_b200 import sharding
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_logical_shards_on_one_gpu(shards):
    n = 200003
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[:] = O.gen_las_fmt0_records(0, n)
    odst = O.OConverter.las_default(ol, olt, scale, offset).convert(osrc, True)
    omn, omx = O.calculate_bounds(odst)
    src = pb.algorithms.synth_las_fmt0_records(n)  # the device generator must equal the oracle's stream
    assert np.array_equal(src.raw_bytes(), osrc.aos[: 20 * n])
    dst = pb.HashMapBuffer(plt, n, "cuda")
    cv = pb.get_default_las_converter(pl, plt, scale, offset)
    merged = sharding.pack_bounds(None, "cuda")
    for r in range(shards):
        rr = sharding.shard_range(n, r, shards)
        part = torch.zeros(6, dtype=torch.float64, device="cuda")
        cv.convert_into_range_with_bounds_device(src, rr, dst, rr, part)
        merged = torch.minimum(merged, part)  # what all_reduce(MIN) does across ranks
    torch.cuda.synchronize()
    mn, mx = sharding.unpack_bounds(merged)
    assert list(mn) == list(omn) and list(mx) == list(omx)
    util.assert_buffers_match(odst, dst)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_bench():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3",
           "--warmup", "3", "--points", "20000000", "--no-e2e", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["scaling"] == "weak"


# ---- peer-memory communicator: convert + fused AABB + exchange in one kernel ------------------------------------------
def _oracle_c2(n):
    ol, pl = util.las_layouts(0, True)
    olt, plt = util.las_layouts(0, False)
    scale, offset = (0.001,) * 3, (500000.0, 5400000.0, 100.0)
    osrc = O.OBuffer(ol, n, False)
    osrc.aos[:] = O.gen_las_fmt0_records(0, n)
    odst = O.OConverter.las_default(ol, olt, scale, offset).convert(osrc, True)
    return pl, plt, scale, offset, odst


@pytest.mark.parametrize("world,empty_rank,direct", [(1, -1, False), (2, -1, False), (3, 1, False), (2, -1, True), (4, -1, False)])
def test_peer_comm_logical_ranks_on_one_gpu(world, empty_rank, direct):
    """W ranks = W contexts with their own streams on ONE GPU, wired by raw pointers; the last CTA of every rank's
    convert kernel publishes its keys into all peers' buffers and waits for theirs.  Five epochs exercise the slot /
    counter parity reuse; an empty rank and the direct kernel take the stand-alone one-CTA path."""
    from pasture_b200.context import Context
    n = 150001
    pl, plt, scale, offset, odst = _oracle_c2(n)
    omn, omx = O.calculate_bounds(odst)
    src = pb.algorithms.synth_las_fmt0_records(n)
    ctxs = [Context(0, use_torch_stream=False) for _ in range(world)]
    if direct:
        for c in ctxs:
            c.set_param("convert.force_direct", 1)
    comms = sharding.PeerComm.local_group(ctxs)
    cvs = [pb.get_default_las_converter(pl, plt, scale, offset, ctx=c) for c in ctxs]
    ranges = []
    for r in range(world):
        rr = sharding.shard_range(n, r, world)
        ranges.append(range(rr.start, rr.start) if r == empty_rank else rr)
    covered = [r for r in ranges if len(r)]
    for epoch in range(5):
        dst = pb.HashMapBuffer(plt, n, "cuda")
        outs = [torch.zeros(6, dtype=torch.float64, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):  # all launches are queued before anything is waited for
            cvs[r].convert_into_range_with_global_bounds(src, ranges[r], dst, ranges[r], comms[r], outs[r])
        for cm in comms:
            cm.check()
        sel = np.concatenate([np.arange(r.start, r.stop) for r in covered])
        pos = odst.attribute_bytes(0).view(np.float64).reshape(-1, 3)[sel]
        for r in range(world):
            mn, mx = sharding.unpack_bounds(outs[r])
            assert list(mn) == list(pos.min(0)) and list(mx) == list(pos.max(0)), (epoch, r)
        if empty_rank < 0:
            util.assert_buffers_match(odst, dst)
            assert list(sharding.unpack_bounds(outs[0])[0]) == list(omn)
    for cm in comms:
        cm.close()
    del cvs  # converters refer to their context: release them first
    import gc
    gc.collect()
    for c in ctxs:
        c.close()  # streams are a finite resource for the logical-rank emulation (see tests/conftest.py)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_peer_comm_bench():
    """real 2-process run: CUDA-IPC mapped exchange buffers over NVLink (bench.py checks the bounds itself)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29578", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "5",
           "--warmup", "3", "--points", "20000000", "--no-e2e", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["n_gpus"] == 2 and d["config"]["bounds_exchange"].startswith("peer memory"), d["config"]
