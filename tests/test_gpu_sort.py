"""GPU parity of the one-sweep radix sort (K8) against torch.sort(stable=True) on the same keys: bit-exact order,
stability of the payload, every bit range / pass count, partial last tiles, skewed and constant digits."""
import numpy as np
import pytest
import torch

from pasture_b200.algorithms import radix_sort

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["digits_8_9", "digits_8"])
def digit_plan(request):
    """9-bit digits are used wherever they save a pass; every sort test also runs with 8-bit digits only"""
    import pasture_b200 as pb
    ctx = pb.get_context()
    ctx.set_param("sort.force_8bit", 1 if request.param == "digits_8" else 0)
    yield request.param
    ctx.set_param("sort.force_8bit", 0)


def _reference(keys, begin, end):
    """stable order by the selected bits, as unsigned"""
    width = end - begin
    sel = (keys >> begin) & ((1 << width) - 1) if width < 64 else keys
    if width == 64:  # unsigned order of int64 bit patterns: flip the sign bit
        sel = sel ^ torch.tensor(-(1 << 63), dtype=torch.int64, device=keys.device)
    return torch.sort(sel, stable=True).indices


@pytest.mark.parametrize("n", [2, 31, 8191, 8192, 8193, 100_003, 1_000_000])
@pytest.mark.parametrize("begin,end", [(0, 64), (0, 8), (27, 61), (0, 63), (5, 14), (40, 41), (3, 12), (10, 36), (1, 19)])
def test_sort_matches_torch_stable_sort(n, begin, end):
    g = torch.Generator(device="cuda").manual_seed(n * 131 + begin * 7 + end)
    keys = torch.randint(-(1 << 63), (1 << 63) - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    order = _reference(keys, begin, end)
    k1 = keys.clone()
    radix_sort(k1, None, begin, end)
    assert torch.equal(k1, keys[order])
    k2, v2 = keys.clone(), torch.arange(n, dtype=torch.int32, device="cuda")
    radix_sort(k2, v2, begin, end)
    assert torch.equal(k2, keys[order]) and torch.equal(v2.long(), order)  # payload = stable permutation


@pytest.mark.parametrize("kind", ["constant", "two_values", "all_ones", "sorted", "reversed", "low_entropy"])
def test_sort_degenerate_distributions(kind):
    n = 300_007
    if kind == "constant":
        keys = torch.full((n,), 0x1234_5678_9ABC, dtype=torch.int64, device="cuda")
    elif kind == "two_values":
        keys = (torch.arange(n, device="cuda") % 2) * 0x0100_0000_0000 + 7
    elif kind == "all_ones":
        keys = torch.full((n,), -1, dtype=torch.int64, device="cuda")  # equals the tile padding pattern
    elif kind == "sorted":
        keys = torch.arange(n, dtype=torch.int64, device="cuda") * 977
    elif kind == "reversed":
        keys = (n - torch.arange(n, dtype=torch.int64, device="cuda")) * 977
    else:
        keys = torch.randint(0, 3, (n,), dtype=torch.int64, device="cuda") << 33
    keys = keys.contiguous()
    order = _reference(keys, 0, 64)
    k, v = keys.clone(), torch.arange(n, dtype=torch.int32, device="cuda")
    radix_sort(k, v, 0, 64)
    assert torch.equal(k, keys[order]) and torch.equal(v.long(), order)


def test_sort_c3_shape_packed_voxel_keys():
    """the voxel grid's use: 34 key bits above 27 index bits, keys only, 20 M keys"""
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    vox = torch.randint(0, 1 << 34, (n,), dtype=torch.int64, device="cuda", generator=g)
    keys = (vox << 27) | torch.arange(n, dtype=torch.int64, device="cuda")
    k = keys.clone()
    radix_sort(k, None, 27, 61)
    assert torch.equal(k, torch.sort(keys).values)  # index in the low bits makes the full-key order the stable order
