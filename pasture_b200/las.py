"""LAS point-block ingest / egress -- host-side mirror of the point path of pasture-io/src/las
(RawLASReader::read_into, raw_readers.rs:366-383; RawLASWriter::write_points_default_layout, raw_writers.rs:203-362).
File handling (VLRs, LAZ, header writing) stays with the caller; this module moves points."""
import ctypes as C

import numpy as np
import torch

from ._lib import LasHeader, LasWriteStats, check, lib
from .context import context_for, get_context
from .layout import PointLayout


def parse_header(file_bytes):
    """-> LasHeader (point_format, record_length, offset_to_point_data, number_of_points, scale, offset, ...)"""
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    h = LasHeader()
    check(lib().pb200_las_parse_header(C.c_void_p(buf.ctypes.data), buf.size, C.byref(h)))
    return h


def default_point_layout(file_bytes):
    """RawLASReader::get_default_point_layout for a file without described extra bytes"""
    return PointLayout.las_default(parse_header(file_bytes).point_format)


def read_points(file_bytes, point_buffer, count=None, first_point=0, buffer_offset=0, ctx=None):
    """PointReader::read_into: fills point_buffer[buffer_offset ...] (any layout, host or device) from the LAS image"""
    ctx = context_for(ctx, point_buffer)
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    h = parse_header(file_bytes)
    n = int(h.number_of_points) - first_point if count is None else int(count)
    d = point_buffer.desc()
    check(lib().pb200_las_read_points(ctx._h, C.c_void_p(buf.ctypes.data), buf.size, first_point, n, C.byref(d), buffer_offset))
    return n


def write_points(point_buffer, point_format, scale, offset, point_range=None, device=None, ctx=None):
    """RawLASWriter::write (default layout): -> (records: uint8 tensor (n, record_length), stats dict)"""
    ctx = context_for(ctx, point_buffer)
    r = point_range if point_range is not None else range(0, point_buffer.len())
    n = len(r)
    rec = PointLayout.las_raw(point_format).size_of_point_entry()
    dev = torch.device(device) if device is not None else point_buffer.device
    out = torch.empty(max(1, n * rec), dtype=torch.uint8, device=dev)  # the writer produces every byte of every record
    st = LasWriteStats()
    d = point_buffer.desc()
    check(lib().pb200_las_write_points(ctx._h, C.byref(d), r.start, r.stop, point_format, (C.c_double * 3)(*scale),
                                       (C.c_double * 3)(*offset), C.c_void_p(out.data_ptr()), 1 if dev.type == "cuda" else 0,
                                       C.byref(st)))
    stats = {"out_of_range": int(st.out_of_range), "points_by_return": [int(x) for x in st.points_by_return],
             "bounds": (tuple(st.bounds_min), tuple(st.bounds_max)) if st.has_bounds else None}
    return out[: n * rec].view(n, rec), stats
