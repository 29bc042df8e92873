"""flow2d-b200: Blackwell-native 2D variational optical flow behind cuda-flow2d's solver surface.

The product is the C-ABI library ``lib/libflow2d_b200.so`` (include/flow2d.h) plus the C++ host
layer in ``host/``.  This Python package is only the ctypes loader used by the tests and bench.py:
it holds no algorithm and has no CPU fallback -- if the library is missing or there is no sm_100
device every compute call raises.

The directory name contains a hyphen; import it through ``flow2d_loader.load()`` at the repo root
(module name ``cuda_flow2d_b200``).
"""
from .binding import (  # noqa: F401
    GREY, GRADIENT, JACOBI, RED_BLACK, TERM_DEFAULT, TERM_GRADIENT, TERM_LOG_GRADIENT, TERM_COMBINED, Flow2D, Flow2DError, Params, build, default_params, level_geometry, level_table,
    lib, lib_path, live_handles, max_warp_level, version,
)
