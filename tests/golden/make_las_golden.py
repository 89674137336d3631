"""Extracts the point records of the reference's LAS test fixtures into a small JSON golden file.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_las_golden.py
Source files: pasture-io/resources/test/10_points_format_{0..10}.las and
10_points_with_extra_bytes_format_{0..10}.las (plain uncompressed LAS 1.4; header parsed with struct).
Only the header numbers needed by the conversion path and the raw point-record bytes are kept; the expected
decoded values live in tests/test_oracle_las.py with their citation (pasture-io/src/las/test_util.rs:46-183).
"""
import json
import os
import struct

REF = "/root/reference/pasture-io/resources/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "las_fixtures.json")


def parse(path):
    b = open(path, "rb").read()
    assert b[:4] == b"LASF"
    header_size, offset_to_points, n_vlrs = struct.unpack_from("<HII", b, 94)
    fmt, rec_len = struct.unpack_from("<BH", b, 104)
    legacy_count = struct.unpack_from("<I", b, 107)[0]
    scale = struct.unpack_from("<3d", b, 131)
    offset = struct.unpack_from("<3d", b, 155)
    maxmin = struct.unpack_from("<6d", b, 179)  # maxx minx maxy miny maxz minz
    count = legacy_count
    if count == 0 and header_size >= 375:
        count = struct.unpack_from("<Q", b, 247)[0]
    records = b[offset_to_points: offset_to_points + count * rec_len]
    assert len(records) == count * rec_len
    return {
        "format": fmt & 0x3F, "record_length": rec_len, "count": count, "scale": list(scale),
        "offset": list(offset), "header_max_min": list(maxmin), "n_vlrs": n_vlrs,
        "records_hex": records.hex(),
        "file_hex": b.hex(),  # the whole (tiny) file: header + VLRs + records, for the ingest tests
    }


def main():
    out = {"plain": {}, "extra_bytes": {}}
    for f in range(11):
        out["plain"][str(f)] = parse(os.path.join(REF, f"10_points_format_{f}.las"))
        out["extra_bytes"][str(f)] = parse(os.path.join(REF, f"10_points_with_extra_bytes_format_{f}.las"))
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
