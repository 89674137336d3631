"""Copies the reference's own .pnts fixture (pasture-io/resources/test/points.pnts, 8000 points, POSITION + RGB) into
tests/golden/ and records, next to it, values decoded from it with nothing but struct/numpy (no pasture code):
the expected header fields, the first/last points and checksums of both arrays. Run in the build container:
    python tests/golden/make_pnts_golden.py
"""
import hashlib
import json
import os
import shutil
import struct

import numpy as np

SRC = "/root/reference/pasture-io/resources/test/points.pnts"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    shutil.copyfile(SRC, os.path.join(HERE, "points.pnts"))
    os.chmod(os.path.join(HERE, "points.pnts"), 0o644)
    b = open(SRC, "rb").read()
    magic, version, byte_length, ft_json, ft_bin, bt_json, bt_bin = struct.unpack("<4s6I", b[:28])
    header = json.loads(b[28:28 + ft_json])
    n = header["POINTS_LENGTH"]
    body = 28 + ft_json
    pos = np.frombuffer(b, dtype="<f4", count=3 * n, offset=body + header["POSITION"]["byteOffset"]).reshape(n, 3)
    rgb = np.frombuffer(b, dtype=np.uint8, count=3 * n, offset=body + header["RGB"]["byteOffset"]).reshape(n, 3)
    out = {
        "source": "pasture-io/resources/test/points.pnts",
        "header": {"magic": magic.decode(), "version": version, "byte_length": byte_length, "feature_table_json_byte_length": ft_json,
                   "feature_table_binary_byte_length": ft_bin, "batch_table_json_byte_length": bt_json,
                   "batch_table_binary_byte_length": bt_bin},
        "feature_table": header,
        "points_length": n,
        "position_sha256": hashlib.sha256(pos.tobytes()).hexdigest(),
        "rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest(),
        "first_positions": pos[:4].tolist(), "last_positions": pos[-4:].tolist(),
        "first_rgb": rgb[:4].tolist(), "last_rgb": rgb[-4:].tolist(),
        "position_min": pos.min(axis=0).tolist(), "position_max": pos.max(axis=0).tolist(),
    }
    json.dump(out, open(os.path.join(HERE, "pnts_fixture.json"), "w"), indent=1)
    print("wrote points.pnts,", n, "points")


if __name__ == "__main__":
    main()
