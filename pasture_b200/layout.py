"""PointLayout descriptors -- host-side mirror of pasture-core/src/layout/point_layout.rs.

Names follow the reference: PointAttributeDataType (:23), PointAttributeDefinition (:261),
PointAttributeMember (:353), FieldAlignment (:601), PointLayout (:648), `attributes` (:454-598).
All offset/size logic runs in the native library (pb200_layout_*).
"""
import ctypes as C
import enum

from ._lib import Attr, check, lib


class PointAttributeDataType(enum.IntEnum):
    U8 = 0
    I8 = 1
    U16 = 2
    I16 = 3
    U32 = 4
    I32 = 5
    U64 = 6
    I64 = 7
    F32 = 8
    F64 = 9
    Vec3u8 = 10
    Vec3u16 = 11
    Vec3f32 = 12
    Vec3i32 = 13
    Vec3f64 = 14
    Vec4u8 = 15
    ByteArray = 16
    Custom = 17

    def size(self, extra_size=0):
        return lib().pb200_dtype_size(int(self), extra_size)

    def min_alignment(self, extra_align=0):
        return lib().pb200_dtype_min_alignment(int(self), extra_align)


DT = PointAttributeDataType


class PointAttributeDefinition:
    """name + datatype (point_layout.rs:261-341)."""

    __slots__ = ("_name", "_datatype", "_extra_size", "_extra_align")

    def __init__(self, name, datatype, extra_size=0, extra_align=0):
        self._name, self._datatype = str(name), DT(datatype)
        self._extra_size, self._extra_align = int(extra_size), int(extra_align)

    @staticmethod
    def custom(name, datatype, extra_size=0, extra_align=0):
        return PointAttributeDefinition(name, datatype, extra_size, extra_align)

    def name(self):
        return self._name

    def datatype(self):
        return self._datatype

    def size(self):
        return self._datatype.size(self._extra_size)

    def with_custom_datatype(self, new_datatype):
        return PointAttributeDefinition(self._name, new_datatype)

    def at_offset_in_type(self, offset):
        return PointAttributeMember(self, offset)

    def __eq__(self, other):
        return (isinstance(other, PointAttributeDefinition) and self._name == other._name
                and self._datatype == other._datatype and self._extra_size == other._extra_size)

    def __hash__(self):
        return hash((self._name, self._datatype, self._extra_size))

    def __repr__(self):
        return f"[{self._name};{self._datatype.name}]"


class PointAttributeMember:
    """attribute + byte offset inside the point record (point_layout.rs:353-431)."""

    __slots__ = ("_definition", "_offset")

    def __init__(self, definition, offset):
        self._definition, self._offset = definition, int(offset)

    @staticmethod
    def custom(name, datatype, offset):
        return PointAttributeMember(PointAttributeDefinition(name, datatype), offset)

    def name(self):
        return self._definition.name()

    def datatype(self):
        return self._definition.datatype()

    def offset(self):
        return self._offset

    def size(self):
        return self._definition.size()

    def attribute_definition(self):
        return self._definition

    def byte_range_within_point(self):
        return range(self._offset, self._offset + self.size())

    def __eq__(self, other):
        return (isinstance(other, PointAttributeMember) and self._definition == other._definition
                and self._offset == other._offset)

    def __repr__(self):
        return f"[{self.name()};{self.datatype().name} @ offset {self._offset}]"


class FieldAlignment:
    """FieldAlignment::Default / FieldAlignment::Packed(n) (point_layout.rs:601-606)."""

    def __init__(self, packed=0):
        self.packed = int(packed)

    @staticmethod
    def Packed(n):
        return FieldAlignment(n)


FieldAlignment.Default = FieldAlignment(0)


class attributes:
    """Built-in attribute definitions (point_layout.rs:454-598)."""
    POSITION_3D = PointAttributeDefinition("Position3D", DT.Vec3f64)
    INTENSITY = PointAttributeDefinition("Intensity", DT.U16)
    RETURN_NUMBER = PointAttributeDefinition("ReturnNumber", DT.U8)
    NUMBER_OF_RETURNS = PointAttributeDefinition("NumberOfReturns", DT.U8)
    CLASSIFICATION_FLAGS = PointAttributeDefinition("ClassificationFlags", DT.U8)
    SCANNER_CHANNEL = PointAttributeDefinition("ScannerChannel", DT.U8)
    SCAN_DIRECTION_FLAG = PointAttributeDefinition("ScanDirectionFlag", DT.U8)
    EDGE_OF_FLIGHT_LINE = PointAttributeDefinition("EdgeOfFlightLine", DT.U8)
    CLASSIFICATION = PointAttributeDefinition("Classification", DT.U8)
    SCAN_ANGLE_RANK = PointAttributeDefinition("ScanAngleRank", DT.I8)
    SCAN_ANGLE = PointAttributeDefinition("ScanAngle", DT.I16)
    USER_DATA = PointAttributeDefinition("UserData", DT.U8)
    POINT_SOURCE_ID = PointAttributeDefinition("PointSourceID", DT.U16)
    COLOR_RGB = PointAttributeDefinition("ColorRGB", DT.Vec3u16)
    GPS_TIME = PointAttributeDefinition("GpsTime", DT.F64)
    NIR = PointAttributeDefinition("NIR", DT.U16)
    WAVE_PACKET_DESCRIPTOR_INDEX = PointAttributeDefinition("WavePacketDescriptorIndex", DT.U8)
    WAVEFORM_DATA_OFFSET = PointAttributeDefinition("WaveformDataOffset", DT.U64)
    WAVEFORM_PACKET_SIZE = PointAttributeDefinition("WaveformPacketSize", DT.U32)
    RETURN_POINT_WAVEFORM_LOCATION = PointAttributeDefinition("ReturnPointWaveformLocation", DT.F32)
    WAVEFORM_PARAMETERS = PointAttributeDefinition("WaveformParameters", DT.Vec3f32)
    POINT_ID = PointAttributeDefinition("PointID", DT.U64)
    NORMAL = PointAttributeDefinition("Normal", DT.Vec3f32)


# pasture-io/src/las/las_layout.rs:37-48
ATTRIBUTE_BASIC_FLAGS = PointAttributeDefinition("LASBasicFlags", DT.U8)
ATTRIBUTE_EXTENDED_FLAGS = PointAttributeDefinition("LASExtendedFlags", DT.U16)
ATTRIBUTE_LOCAL_LAS_POSITION = PointAttributeDefinition("LASLocalPosition", DT.Vec3i32)


class PointLayout:
    """point_layout.rs:648-997 over a native pb200_layout handle."""

    def __init__(self, _handle=None):
        if _handle is None:
            h = C.c_void_p()
            check(lib().pb200_layout_create(C.byref(h)))
            _handle = h
        self._h = _handle

    def __del__(self):
        try:
            if self._h:
                lib().pb200_layout_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- constructors ------------------------------------------------------------------------------
    @classmethod
    def default(cls):
        return cls()

    @classmethod
    def from_attributes(cls, attrs):  # :669
        l = cls()
        for a in attrs:
            l.add_attribute(a, FieldAlignment.Default)
        return l

    @classmethod
    def from_attributes_packed(cls, attrs, max_alignment):  # :693
        l = cls()
        for a in attrs:
            l.add_attribute(a, FieldAlignment.Packed(max_alignment))
        return l

    @classmethod
    def from_members_and_alignment(cls, members, type_alignment):  # :719
        arr = (Attr * max(1, len(members)))()
        for i, m in enumerate(members):
            arr[i].name = m.name().encode()
            arr[i].dtype = int(m.datatype())
            arr[i].extra_size = m.attribute_definition()._extra_size
            arr[i].extra_align = m.attribute_definition()._extra_align
            arr[i].offset = m.offset()
        h = C.c_void_p()
        check(lib().pb200_layout_from_members_and_alignment(arr, len(members), type_alignment, C.byref(h)))
        return cls(h)

    @classmethod
    def las_raw(cls, point_format):
        """point_layout_from_las_point_format(format, exact_binary_representation=true), las_layout.rs:70-108"""
        h = C.c_void_p()
        check(lib().pb200_las_raw_layout(point_format, C.byref(h)))
        return cls(h)

    @classmethod
    def las_default(cls, point_format):
        """LasPointFormatN::layout(), las_types.rs"""
        h = C.c_void_p()
        check(lib().pb200_las_default_layout(point_format, C.byref(h)))
        return cls(h)

    def clone(self):
        h = C.c_void_p()
        check(lib().pb200_layout_clone(self._h, C.byref(h)))
        return PointLayout(h)

    # -- mutation ----------------------------------------------------------------------------------
    def add_attribute(self, point_attribute, field_alignment=None):  # :778
        fa = field_alignment or FieldAlignment.Default
        check(lib().pb200_layout_add_attribute(self._h, point_attribute.name().encode(), int(point_attribute.datatype()),
                                               point_attribute._extra_size, point_attribute._extra_align, fa.packed))

    # -- queries -----------------------------------------------------------------------------------
    def __len__(self):
        return lib().pb200_layout_num_attributes(self._h)

    def at(self, index):  # :898
        a = Attr()
        check(lib().pb200_layout_get_attribute(self._h, index, C.byref(a)))
        d = PointAttributeDefinition(a.name.decode(), a.dtype, a.extra_size, a.extra_align)
        return PointAttributeMember(d, a.offset)

    def attributes(self):  # :914
        return [self.at(i) for i in range(len(self))]

    def size_of_point_entry(self):  # :930
        return lib().pb200_layout_size_of_point_entry(self._h)

    def alignment(self):
        return lib().pb200_layout_alignment(self._h)

    def has_attribute_with_name(self, name):  # :831
        return lib().pb200_layout_index_by_name(self._h, name.encode()) >= 0

    def has_attribute(self, attribute):  # :850
        return lib().pb200_layout_index_of(self._h, attribute.name().encode(), int(attribute.datatype())) >= 0

    def get_attribute(self, attribute):  # :868
        i = lib().pb200_layout_index_of(self._h, attribute.name().encode(), int(attribute.datatype()))
        return self.at(i) if i >= 0 else None

    def get_attribute_by_name(self, name):  # :887
        i = lib().pb200_layout_index_by_name(self._h, name.encode())
        return self.at(i) if i >= 0 else None

    def index_by_name(self, name):
        i = lib().pb200_layout_index_by_name(self._h, name.encode())
        return i if i >= 0 else None

    def index_of(self, attribute):  # :951
        i = lib().pb200_layout_index_of(self._h, attribute.name().encode(), int(attribute.datatype()))
        return i if i >= 0 else None

    def offset_of(self, attribute):  # :975
        m = self.get_attribute(attribute)
        return m.offset() if m is not None else None

    def compare_without_offsets(self, other):  # :960
        return bool(lib().pb200_layout_compare_without_offsets(self._h, other._h))

    def __eq__(self, other):
        return isinstance(other, PointLayout) and bool(lib().pb200_layout_equal(self._h, other._h))

    def __repr__(self):
        return "PointLayout {\n" + "".join(f"\t{a}\n" for a in self.attributes()) + "}"
