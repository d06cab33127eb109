"""ps3d_b200 — B200-native implementation of the ps3d time-step hot path.

The product is `csrc/` (hand-written sm_100a kernels + the C ABI declared in
`include/ps3d_cuda.h`), built in-tree as `ps3d_b200/libps3d_cuda.so`.  This
package is only the Python binding used by the tests and `bench.py`; the
Fortran host program binds the same C ABI through `iso_c_binding`
(INTEGRATION.md).  There is no CPU fallback: loading fails loudly when the
shared library is missing, and `init` fails when no CUDA device is present.
"""
from .lib import PS3DLib, PS3DError, load, LIB_PATH  # noqa: F401
from . import host  # noqa: F401
