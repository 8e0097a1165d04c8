"""Fiber scheduler behind the one-call hot-path API (smalt_b200/hostc/shim_fiber.inc.c): the
CPU-only self test - every item runs to completion on its own stack, parked calls are released
in waves, items that ask for their turn (random draws, results.c:2298) get it in item order."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smalt_b200", "libsmalt_b200_map.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="driver library not built")
@pytest.mark.parametrize("nfib,nitems", [(1, 10), (4, 50), (64, 1000), (512, 5000), (8, 0), (2048, 300)])
def test_fiber_scheduler(nfib, nitems):
    lib = C.CDLL(LIB, mode=os.RTLD_LOCAL)
    order = np.full(max(nitems, 1), -1, np.int32)
    draws = np.full(max(nitems, 1), -1, np.int32)
    nd = C.c_int(0)
    rc = lib.smbFiberSelfTest(nfib, nitems, order.ctypes.data_as(C.c_void_p), draws.ctypes.data_as(C.c_void_p),
                              C.byref(nd))
    assert rc == nitems
    assert sorted(order[:nitems].tolist()) == list(range(nitems))       # every item ran exactly once
    d = draws[:nd.value]
    assert nd.value == len(range(0, nitems, 3)) and (np.diff(d) > 0).all()  # draws strictly in item order
    if nfib > 1 and nitems > 4 * nfib:
        assert (order[:nitems] != np.arange(nitems)).any()               # items really interleave
