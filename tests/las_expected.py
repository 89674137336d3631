"""Expected decoded values of the reference's LAS fixtures (pasture-io/src/las/test_util.rs:46-183)."""
import numpy as np

N = 10
POSITIONS = np.array([[i, i, i] for i in range(10)], dtype=np.float64)            # :50-63
INTENSITIES = np.array([i * 255 for i in range(10)], dtype=np.uint16)             # :65-78
RETURN_NUMBERS = np.array([0, 1, 2, 3, 4, 5, 6, 7, 0, 1], dtype=np.uint8)         # :80-82
RETURN_NUMBERS_EXT = np.arange(10, dtype=np.uint8)                                # :84-86
NUMBER_OF_RETURNS = RETURN_NUMBERS.copy()                                         # :88-90
NUMBER_OF_RETURNS_EXT = np.arange(10, dtype=np.uint8)                             # :92-94
CLASSIFICATION_FLAGS = np.arange(10, dtype=np.uint8)                              # :96-98
SCANNER_CHANNELS = np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1], dtype=np.uint8)       # :100-102
SCAN_DIRECTION_FLAGS = np.array([0, 1] * 5, dtype=np.uint8)                       # :104-106
EDGE_OF_FLIGHT_LINES = np.array([0, 1] * 5, dtype=np.uint8)                       # :108-110
CLASSIFICATIONS = np.arange(10, dtype=np.uint8)                                   # :112-114
SCAN_ANGLE_RANKS = np.arange(10, dtype=np.int8)                                   # :116-118
SCAN_ANGLES_EXT = np.arange(10, dtype=np.int16)                                   # :120-122
USER_DATA = np.arange(10, dtype=np.uint8)                                         # :124-126
POINT_SOURCE_IDS = np.arange(10, dtype=np.uint16)                                 # :128-130
GPS_TIMES = np.arange(1, 11, dtype=np.float64)                                    # :132-134
COLORS = np.array([[i, (i + 1) << 4, (i + 2) << 8] for i in range(10)], dtype=np.uint16)  # :136-149
NIRS = np.arange(10, dtype=np.uint16)                                             # :151-153
WAVEPACKET_INDEX = np.arange(10, dtype=np.uint8)                                  # :155-157
WAVEPACKET_OFFSET = np.arange(10, dtype=np.uint64)                                # :159-161
WAVEPACKET_SIZE = np.arange(10, dtype=np.uint32)                                  # :163-165
WAVEPACKET_LOCATION = np.arange(10, dtype=np.float32)                             # :167-169
WAVEPACKET_PARAMETERS = np.array([[i + 1, i + 2, i + 3] for i in range(10)], dtype=np.float32)  # :171-183


def fmt_flags(f):
    return dict(extended=f >= 6, gps=f in (1, 3, 4, 5, 6, 7, 8, 9, 10), color=f in (2, 3, 5, 7, 8, 10),
                nir=f in (8, 10), waveform=f in (4, 5, 9, 10))


def expected_default_layout_values(f):
    """attribute name -> expected array for the default (non-raw) layout of LAS format f
    (compare_to_reference_data_range, test_util.rs:186-420)"""
    fl = fmt_flags(f)
    e = {"Position3D": POSITIONS, "Intensity": INTENSITIES,
         "ReturnNumber": RETURN_NUMBERS_EXT if fl["extended"] else RETURN_NUMBERS,
         "NumberOfReturns": NUMBER_OF_RETURNS_EXT if fl["extended"] else NUMBER_OF_RETURNS,
         "ScanDirectionFlag": SCAN_DIRECTION_FLAGS, "EdgeOfFlightLine": EDGE_OF_FLIGHT_LINES,
         "Classification": CLASSIFICATIONS, "UserData": USER_DATA, "PointSourceID": POINT_SOURCE_IDS}
    if fl["extended"]:
        e["ClassificationFlags"] = CLASSIFICATION_FLAGS
        e["ScannerChannel"] = SCANNER_CHANNELS
        e["ScanAngle"] = SCAN_ANGLES_EXT
    else:
        e["ScanAngleRank"] = SCAN_ANGLE_RANKS
    if fl["gps"]:
        e["GpsTime"] = GPS_TIMES
    if fl["color"]:
        e["ColorRGB"] = COLORS
    if fl["nir"]:
        e["NIR"] = NIRS
    if fl["waveform"]:
        e["WavePacketDescriptorIndex"] = WAVEPACKET_INDEX
        e["WaveformDataOffset"] = WAVEPACKET_OFFSET
        e["WaveformPacketSize"] = WAVEPACKET_SIZE
        e["ReturnPointWaveformLocation"] = WAVEPACKET_LOCATION
        e["WaveformParameters"] = WAVEPACKET_PARAMETERS
    return e
