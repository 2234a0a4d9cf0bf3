"""pyflwdir_b200 -- B200-native (sm_100a CUDA) implementation of pyflwdir's D8 flow-network hot path behind the
reference's own `from_array` / `FlwdirRaster` API.

Host code is plain Python + numpy calling hand-written CUDA through a C ABI (ctypes, `include/pfd_b200.h`); no
PyTorch, no Triton, no CPU fallback: without the built `libpfd_b200.so` and a CUDA device every compute call
raises. See DESIGN.md for the kernels and INTEGRATION.md for the drop-in boundary.
"""
from . import arithmetics, basins, core, core_conversion, core_d8, core_ldd, core_nextxy, dem, geotiff, gis_utils, regions, rivers, streams, subgrid, upscale
from .core_nextxy import read_nextxy
from .core_conversion import d8_to_ldd, ldd_to_d8
from .gis_utils import Affine
from .pyflwdir import FlwdirRaster, from_array, from_dem, _get_idxs_dtype
from .flwdir import Flwdir

__version__ = "0.1.0"
__all__ = ["FlwdirRaster", "Flwdir", "from_array", "from_dem", "gis_utils", "Affine", "d8_to_ldd", "ldd_to_d8", "read_nextxy", "geotiff"]
