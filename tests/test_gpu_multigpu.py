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
