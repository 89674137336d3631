"""Shared helpers for the .pnts tests: the committed reference fixture and its independently decoded expectations."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture():
    blob = open(os.path.join(GOLDEN, "points.pnts"), "rb").read()
    exp = json.load(open(os.path.join(GOLDEN, "pnts_fixture.json")))
    return blob, exp


def check_fixture_arrays(pos, rgb, exp, rtc=None):
    pos = np.ascontiguousarray(pos)
    rgb = np.ascontiguousarray(rgb)
    assert pos.shape == (exp["points_length"], 3) and rgb.shape == (exp["points_length"], 3)
    if rtc is None:
        assert hashlib.sha256(pos.astype("<f4").tobytes()).hexdigest() == exp["position_sha256"]
        assert pos[:4].tolist() == exp["first_positions"] and pos[-4:].tolist() == exp["last_positions"]
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == exp["rgb_sha256"]
    assert rgb[:4].tolist() == exp["first_rgb"] and rgb[-4:].tolist() == exp["last_rgb"]
