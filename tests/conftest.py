import os
import sys

import os

# several spin-waiting kernels of ONE process on ONE device (logical ranks in test_gpu_multigpu.py) must not share a
# hardware work queue: with the default 8 connections, streams alias and a waiting kernel can block the kernel it waits for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def las_fixtures():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "las_fixtures.json")) as fh:
        return json.load(fh)
