"""smalt_b200 - B200-native (sm_100a CUDA) hot path of the SMALT read mapper.

The product is the C-ABI shared library ``libsmalt_b200.so`` (include/smalt_b200.h,
sources in smalt_b200/csrc).  This package is the thin Python host-side binding used
by the tests and by bench.py; it loads the library with ctypes and fails loudly if
it has not been built (there is no CPU fallback)."""
from .capi import Context, SmbError, lib_path, load_library  # noqa: F401
