"""granges_b200 -- B200-native (sm_100a CUDA) overlap-join hot path of vsbuffalo/granges.

The product is the C-ABI shared library `libgranges_b200.so` (include/granges_b200.h); this package
is its Python face.  Importing it requires the built library -- there is no CPU fallback.
"""
from . import _capi
from .api import OPS, Context, GRangesError, Index, PreparedMap, Windows, map_whole_mgpu, shard_plan

from . import granges  # noqa: E402  (the reference's container names over the C ABI)

_capi.lib()  # fail loudly at import time if the CUDA library is missing

__all__ = ["Context", "Index", "Windows", "PreparedMap", "GRangesError", "OPS", "shard_plan", "map_whole_mgpu"]
