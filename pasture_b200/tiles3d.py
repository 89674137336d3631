"""3D-Tiles .pnts reader / writer -- host-side mirror of pasture-io/src/tiles3d (PntsReader pnts_reader.rs:41-403,
PntsWriter pnts_writer.rs:68-401). The 28-byte header and the JSON FeatureTable header are handled here; the point
data (FeatureTable binary body <-> point buffer, datatype casts, RTC_CENTER) goes through the library."""
import ctypes as C
import json
import struct

import numpy as np
import torch

from ._lib import Attr, check, lib
from .containers import HashMapBuffer, VectorBuffer
from .context import context_for, get_context
from .layout import DT, FieldAlignment, PointAttributeDefinition, PointLayout, attributes

COLOR_RGBA = PointAttributeDefinition("ColorRGBA", DT.Vec4u8)  # pnts_types.rs:11-14
PNTS_HEADER_BYTE_LENGTH = 28  # pnts_types.rs:33
RELATIVE_TO_CENTER, ABSOLUTE = "RelativeToCenter", "Absolute"  # PntsReadPositionsMode, pnts_reader.rs:30-38

# semantic -> (attribute definition as stored in the file), in the order layout_from_feature_table_header checks them
_SEMANTICS = [("POSITION", attributes.POSITION_3D.with_custom_datatype(DT.Vec3f32)), ("RGBA", COLOR_RGBA),
              ("RGB", attributes.COLOR_RGB.with_custom_datatype(DT.Vec3u8)), ("NORMAL", attributes.NORMAL)]
_SEMANTIC_OF = {"Position3D": "POSITION", "ColorRGB": "RGB", "ColorRGBA": "RGBA", "Normal": "NORMAL"}  # pnts_writer.rs:38-52


def _align8(v):
    return (v + 7) & ~7


def _attr_array(entries):
    arr = (Attr * max(1, len(entries)))()
    for i, (name, dtype, offset) in enumerate(entries):
        arr[i].name = name.encode()
        arr[i].dtype = int(dtype)
        arr[i].offset = int(offset)
        arr[i].size = lib().pb200_dtype_size(int(dtype), 0)
    return arr


class PntsMetadata:
    def __init__(self, points_length, rtc_center=None, quantized_volume_offset=None, quantized_volume_scale=None,
                 constant_rgba=None, batch_length=None):
        self.points_length = points_length
        self.rtc_center = rtc_center
        self.quantized_volume_offset = quantized_volume_offset
        self.quantized_volume_scale = quantized_volume_scale
        self.constant_rgba = constant_rgba
        self.batch_length = batch_length

    def number_of_points(self):
        return self.points_length


class PntsReader:
    def __init__(self, file_bytes, device=None):
        """`file_bytes`: the .pnts image (bytes / numpy uint8 / torch uint8 tensor; a CUDA tensor keeps the body in HBM)"""
        if isinstance(file_bytes, torch.Tensor):
            self._blob = file_bytes.contiguous().view(torch.uint8)
            head = bytes(self._blob[: min(self._blob.numel(), 1 << 20)].cpu().numpy())
        else:
            head = bytes(file_bytes)
            self._blob = torch.from_numpy(np.frombuffer(head, dtype=np.uint8).copy())
        if device is not None:
            self._blob = self._blob.to(device)
        magic, _version, _byte_length, ft_json, _ft_bin, _bt_json, _bt_bin = struct.unpack("<4s6I", head[:PNTS_HEADER_BYTE_LENGTH])
        if magic != b"pnts":  # verify_magic, pnts_types.rs:58-63
            raise ValueError(f"No valid PNTS file, expected first four bytes to be equal to 'pnts', but was '{magic}' instead")
        if PNTS_HEADER_BYTE_LENGTH + ft_json > self._blob.numel():
            raise EOFError("unexpected end of file inside the FeatureTable JSON header")
        if PNTS_HEADER_BYTE_LENGTH + ft_json > len(head):  # device-resident image: fetch the whole JSON, however long
            head = bytes(self._blob[: PNTS_HEADER_BYTE_LENGTH + ft_json].cpu().numpy())
        header = json.loads(head[PNTS_HEADER_BYTE_LENGTH:PNTS_HEADER_BYTE_LENGTH + ft_json].decode("utf-8"))
        if not isinstance(header, dict):
            raise ValueError("FeatureTable JSON header was no JSON object")
        body_offset = PNTS_HEADER_BYTE_LENGTH + ft_json  # the reader continues where the JSON ended (feature_table.rs:119)
        self.layout = PointLayout.default()
        self.attribute_offsets = {}
        for semantic, definition in _SEMANTICS:  # pnts_reader.rs:107-176
            if semantic in header:
                ref = header[semantic]
                if not (isinstance(ref, dict) and "byteOffset" in ref):
                    raise ValueError(f"Found PNTS attribute {semantic} ({ref}) but it was not a reference to the feature table binary!")
                self.attribute_offsets[definition.name()] = body_offset + int(ref["byteOffset"])
                self.layout.add_attribute(definition, FieldAlignment.Packed(1))
        if "POINTS_LENGTH" not in header:
            raise ValueError("Mandatory value POINTS_LENGTH not found in feature table header")
        n = header["POINTS_LENGTH"]
        if not isinstance(n, int) or isinstance(n, bool) or n < 0:
            raise ValueError("POINTS_LENGTH value vas no integer number")

        def vec(key, count):
            v = header.get(key)
            if v is None:
                return None
            if not isinstance(v, list) or len(v) != count:
                raise ValueError(f"{key} value was no array entry of length {count}")
            return [float(x) for x in v] if count == 3 else [int(x) for x in v]

        self.metadata = PntsMetadata(n, vec("RTC_CENTER", 3), vec("QUANTIZED_VOLUME_OFFSET", 3), vec("QUANTIZED_VOLUME_SCALE", 3),
                                     vec("CONSTANT_RGBA", 4), header.get("BATCH_LENGTH"))
        self.current_point_index = 0
        self._mode = ABSOLUTE

    def set_read_positions_mode(self, mode):
        assert mode in (ABSOLUTE, RELATIVE_TO_CENTER)
        self._mode = mode

    def read_positions_mode(self):
        return self._mode

    def get_metadata(self):
        return self.metadata

    def get_default_point_layout(self):
        return self.layout

    def read_into(self, point_buffer, count, ctx=None):  # pnts_reader.rs:294-367
        ctx = context_for(ctx, point_buffer)
        remaining = self.metadata.points_length - self.current_point_index
        num_to_read = min(remaining, int(count))
        if num_to_read == 0:
            raise EOFError("No points remaining in PNTS file")
        blob = self._blob if self._blob.device == point_buffer.device else self._blob.to(point_buffer.device)
        entries = [(m.name(), m.datatype(), self.attribute_offsets[m.name()]) for m in self.layout.attributes()]
        arr = _attr_array(entries)
        rtc = None
        if self._mode == ABSOLUTE and self.metadata.rtc_center is not None:
            rtc = (C.c_double * 3)(*self.metadata.rtc_center)
        d = point_buffer.desc()
        check(lib().pb200_pnts_read_points(ctx._h, C.c_void_p(blob.data_ptr()), blob.numel(), arr, len(entries), self.current_point_index,
                                           num_to_read, C.byref(d), rtc))
        self.current_point_index += num_to_read
        return num_to_read

    def read(self, count, buffer_type=VectorBuffer, device="cpu", ctx=None):
        """PointReader::read: a new buffer in the reader's default layout"""
        n = min(self.metadata.points_length - self.current_point_index, int(count))
        buf = buffer_type(self.layout, n, device)
        if n:
            self.read_into(buf, n, ctx)
        return buf

    def seek_point(self, index):  # SeekFrom::Start, pnts_reader.rs:379-397
        self.current_point_index = max(0, min(int(index), self.metadata.points_length))
        return self.current_point_index


class PntsWriter:
    def __init__(self, point_layout, device="cpu"):  # from_write_and_layout, pnts_writer.rs:81-96
        self.expected_layout = point_layout
        self.rtc_center = None
        self._device = torch.device(device)
        self._chunks = []  # (n, body tensor) per write call; merged in flush like the cached HashMapBuffer
        arr = (Attr * 4)()
        k, nbytes = C.c_uint32(0), C.c_uint64(0)
        check(lib().pb200_pnts_compatible_layout(point_layout._h, 0, arr, C.byref(k), C.byref(nbytes)))
        self.default_layout = PointLayout.default()
        self._attrs = []
        for i in range(k.value):
            definition = PointAttributeDefinition(arr[i].name.decode(), DT(arr[i].dtype))
            self.default_layout.add_attribute(definition, FieldAlignment.Default)
            self._attrs.append(definition)

    def set_rtc_center(self, rtc_center):  # :98-103 -- does NOT translate the points
        self.rtc_center = [float(x) for x in rtc_center]

    def get_default_point_layout(self):
        return self.default_layout

    def write(self, points, ctx=None):  # :353-401
        ctx = context_for(ctx, points)
        if points.point_layout() != self.expected_layout:
            raise ValueError("PointLayout of buffer does not match the PointLayout that this PntsWriter was constructed with!")
        n = points.len()
        arr = (Attr * 4)()
        k, nbytes = C.c_uint32(0), C.c_uint64(0)
        check(lib().pb200_pnts_compatible_layout(points.point_layout()._h, n, arr, C.byref(k), C.byref(nbytes)))
        # every array and its padding to 8 bytes is written by the library (pnts.cu: pb200_pnts_write_points)
        body = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=points.device)
        d = points.desc()
        check(lib().pb200_pnts_write_points(ctx._h, C.byref(d), C.c_void_p(body.data_ptr()), nbytes.value))
        self._chunks.append((n, [(int(arr[i].offset), int(arr[i].size)) for i in range(k.value)], body))

    def _feature_table_body(self):
        """one padded array per attribute over all cached points (write_feature_table_body, :310-341)"""
        total = sum(n for n, _, _ in self._chunks)
        parts, offsets, off = [], [], 0
        for i, _ in enumerate(self._attrs):
            offsets.append(off)
            size = 0
            for n, lay, body in self._chunks:
                o, sz = lay[i]
                parts.append(body[o:o + n * sz].cpu())
                size += n * sz
            pad = _align8(size) - size
            if pad:
                parts.append(torch.zeros(pad, dtype=torch.uint8))
            off += _align8(size)
        body = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8)
        return total, offsets, bytes(body.numpy())

    def flush(self):
        """-> the complete .pnts image (write_cached_points, :152-232)"""
        total, offsets, body = self._feature_table_body()
        header = {}
        for definition, off in zip(self._attrs, offsets):
            header[_SEMANTIC_OF[definition.name()]] = {"byteOffset": off}
        header["POINTS_LENGTH"] = total
        if self.rtc_center is not None:
            header["RTC_CENTER"] = self.rtc_center

        def json_header(obj, position_in_file):  # write_json_header, common.rs:29-55: padded with spaces to 8 bytes
            text = json.dumps(obj, separators=(",", ":")).encode()
            return text + b" " * (_align8(position_in_file + len(text)) - (position_in_file + len(text)))

        ft = json_header(header, PNTS_HEADER_BYTE_LENGTH)
        body_aligned = _align8(PNTS_HEADER_BYTE_LENGTH + len(ft) + len(body)) - (PNTS_HEADER_BYTE_LENGTH + len(ft))
        body = body + b"\0" * (body_aligned - len(body))
        start_bt = PNTS_HEADER_BYTE_LENGTH + len(ft) + body_aligned
        bt = json_header({}, start_bt)
        total_len = start_bt + len(bt)
        if total_len >= 1 << 32:
            raise OverflowError("Size of .pnts file exceeds maximum size of 4GiB!")
        head = struct.pack("<4s6I", b"pnts", 1, total_len, len(ft), body_aligned, len(bt), 0)
        return head + ft + body + bt
