"""pasture_b200 -- B200-native (sm_100a) implementation of pasture's per-point hot path.

The product is libpasture_b200.so (hand-written CUDA behind the C ABI of include/pasture_b200.h); this
package is the host-side mirror of the reference's Rust interface for that path (same names, argument
meaning and error behaviour) used by the parity tests and the benchmark.  Importing it loads the shared
library and fails loudly if the library has not been built; there is no CPU fallback.
"""
from . import _lib
from ._lib import PastureB200Error

_lib.lib()  # fail at import time, not at first use

from .layout import (ATTRIBUTE_BASIC_FLAGS, ATTRIBUTE_EXTENDED_FLAGS, ATTRIBUTE_LOCAL_LAS_POSITION,  # noqa: E402
                     FieldAlignment, PointAttributeDataType, PointAttributeDefinition, PointAttributeMember,
                     PointLayout, attributes)
from .containers import HashMapBuffer, VectorBuffer, buffers_equal, filter, filter_into  # noqa: E402
from .context import Context, get_context, kernel_launch_count  # noqa: E402
from .conversion import (Add, BufferLayoutConverter, InvScaleOffset, ScaleOffset, ShiftMask, Transform,  # noqa: E402
                         get_default_las_converter, transform_attribute, view_attribute_with_conversion)
from . import algorithms  # noqa: E402
from . import sharding  # noqa: E402
from . import las  # noqa: E402
from . import tiles3d  # noqa: E402

__all__ = [
    "PastureB200Error", "PointAttributeDataType", "PointAttributeDefinition", "PointAttributeMember", "PointLayout",
    "FieldAlignment", "attributes", "ATTRIBUTE_BASIC_FLAGS", "ATTRIBUTE_EXTENDED_FLAGS",
    "ATTRIBUTE_LOCAL_LAS_POSITION", "VectorBuffer", "HashMapBuffer", "buffers_equal", "filter", "filter_into", "Context", "get_context",
    "kernel_launch_count", "BufferLayoutConverter", "Transform", "ScaleOffset", "InvScaleOffset", "Add", "ShiftMask",
    "get_default_las_converter", "transform_attribute", "view_attribute_with_conversion", "algorithms", "sharding", "las", "tiles3d",
]
